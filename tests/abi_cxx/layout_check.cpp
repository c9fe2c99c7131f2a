// layout_check.cpp -- compile-time proof that the C ABI structs have the layout of the reference's own
// structs.  Includes the reference headers from where they lie (never copied):
//   GPU/detail/format.h  (CsrMatrix, CsrMatrixDev, CooMatrixDev, DiaMatrixDev, EllMatrixDev; needs the CUSP stub)
//   CPU/detail/format.h  (CsrMatrix, CooMatrix; needs oracle/ref_shim/mkl.h)
// Built by tests/test_abi.py::test_layouts_against_the_reference_headers with
//   g++ -std=c++14 -fsyntax-only -DVALUE_TYPE=double -I tests/abi_cxx/cusp_stub -I oracle/ref_shim
//       -DREF_GPU_FORMAT_H=... -DREF_CPU_FORMAT_H=... tests/abi_cxx/layout_check.cpp
// A successful compile is the test.
#include <stddef.h>
#include <stdio.h>

#include "../../include/iaspgemm.h"

#include <cusp/multiply.h>
#include <cusp/array2d.h>
#include <cusp/print.h>
#include "mkl.h"

namespace refgpu {
#include REF_GPU_FORMAT_H
}
#undef FORMAT_H
namespace refcpu {
#include REF_CPU_FORMAT_H
}

#define SAME_FIELD(A, B, fa, fb)                                                                   \
    static_assert(offsetof(A, fa) == offsetof(B, fb), #A "::" #fa " is not where " #B "::" #fb " is"); \
    static_assert(sizeof(((A *)0)->fa) == sizeof(((B *)0)->fb), #A "::" #fa " has another size than " #B "::" #fb)

// ---- CSR (GPU/detail/format.h:47-69, CPU/detail/format.h:29-39)
static_assert(sizeof(IasCsrMatrix) == sizeof(refgpu::CsrMatrix), "CsrMatrix");
static_assert(sizeof(IasCsrMatrix) == sizeof(refcpu::CsrMatrix), "CsrMatrix (CPU)");
static_assert(sizeof(IasCsrMatrixDev) == sizeof(refgpu::CsrMatrixDev), "CsrMatrixDev");
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, choice, choice);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, row, row);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, col, col);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, nnz, nnz);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, row_ind, row_ind);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, col_ind, col_ind);
SAME_FIELD(IasCsrMatrix, refgpu::CsrMatrix, values, values);
SAME_FIELD(IasCsrMatrix, refcpu::CsrMatrix, row_ind, row_ind);
SAME_FIELD(IasCsrMatrix, refcpu::CsrMatrix, values, values);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, choice, choice);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, row, row);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, col, col);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, nnz, nnz);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, row_ind_dev, row_ind_dev);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, col_ind_dev, col_ind_dev);
SAME_FIELD(IasCsrMatrixDev, refgpu::CsrMatrixDev, values_dev, values_dev);
// the cuSPARSE bridge struct is the same shape (GPU/detail/format.h:121-131)
static_assert(sizeof(IasCsrMatrixDev) == sizeof(refgpu::cuSparseMatrix), "cuSparseMatrix");
SAME_FIELD(IasCsrMatrixDev, refgpu::cuSparseMatrix, values_dev, values_dev);

// ---- COO (GPU/detail/format.h:29-40)
static_assert(sizeof(IasCooDev) == sizeof(refgpu::CooMatrixDev), "CooMatrixDev");
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, choice, choice);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, row, row);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, col, col);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, nnz, nnz);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, row_offset_dev, row_offset_dev);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, row_ind_dev, row_ind_dev);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, col_ind_dev, col_ind_dev);
SAME_FIELD(IasCooDev, refgpu::CooMatrixDev, values_dev, values_dev);
static_assert(sizeof(IasCooDev) == sizeof(refcpu::CooMatrix), "CooMatrix (CPU, host twin)");

// ---- DIA (GPU/detail/format.h:82-92)
static_assert(sizeof(IasDiaDev) == sizeof(refgpu::DiaMatrixDev), "DiaMatrixDev");
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, choice, choice);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, row, row);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, col, col);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, num_diagonals, num_diagonals);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, diagonal_ind_dev, diagonal_ind_dev);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, diagonal_offsets_dev, diagonal_offsets_dev);
SAME_FIELD(IasDiaDev, refgpu::DiaMatrixDev, values_dev, values_dev);

// ---- ELL (GPU/detail/format.h:108-119)
static_assert(sizeof(IasEllDev) == sizeof(refgpu::EllMatrixDev), "EllMatrixDev");
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, choice, choice);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, row, row);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, col, col);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, nnz, nnz);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, max_nnz_per_row, max_nnz_per_row);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, nnz_row_dev, nnz_row_dev);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, col_ind_dev, col_ind_dev);
SAME_FIELD(IasEllDev, refgpu::EllMatrixDev, values_dev, values_dev);

int main() { return 0; }
