// auto_api.cu -- the front end's path as one call on host operands: features -> format selection -> conversion ->
// multiply in the selected format -> result on the host.
//
// Mirrors what the reference's main does between loading the .mtx files and the report block
// (CPU/main.cpp:655-704 features + MatNet.Pred, :658-676 conversions, :746-935 the multiplies), except that the
// selected algorithm is the one that runs (the reference predicts and then runs everything, Appendix D of SURVEY.md).
// The result comes back in the selected format's own layout, as every reference kernel returns it:
// CSR (CSR_MUL_CSR, csr:85), DIA row-major values[row][diag] (DIA_mul_DIA, dia:101), ELL row-major (ELL_MUL_ELL, ell:80).
#include <time.h>

#include <algorithm>

#include "common.cuh"

using namespace ias;

namespace {

inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

double ms_between(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

}  // namespace

extern "C" {

// The rule the front end falls back on when no MatNet weights are at hand: class index in the CPU numbering
// (1 = CSR -- the reference's slots 0/1 are MKL/CSR --, 2 = DIA, 3 = ELL).  f = the 26 features (CPU/main.cpp:655-679).
int ias_select_format(const double *f, int dia_ok, int ell_ok)
{
    if (!f) return 1;
    const double diag_fill = (f[18] > 0 && f[0] > 0) ? f[2] / (f[18] * f[0]) : 0.0;      // nnz / (ndiag * rows)
    const double ell_eff = f[24];
    if (dia_ok && diag_fill > 0.5) return 2;
    if (ell_ok && ell_eff > 0.9 && f[8] < 0.05) return 3;
    return 1;
}

int ias_spgemm_auto_host(const IasCsrMatrix *A, const IasCsrMatrix *B, double gate, void *matnet, IasAutoResult *out)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !out) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    memset(out, 0, sizeof *out);
    struct timespec t_in, t_out;
    clock_gettime(CLOCK_MONOTONIC, &t_in);
    auto stamp = [&](int k) {
        struct timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        out->ms_host[k] = (t.tv_sec - t_in.tv_sec) * 1e3 + (t.tv_nsec - t_in.tv_nsec) / 1e6;
    };
    const bool alias = (A == B) || (A->row_ind == B->row_ind && A->col_ind == B->col_ind && A->values == B->values &&
                                    A->row == B->row && A->col == B->col);
    cudaEvent_t *ev = c.ev_bin + 24;             // eight events nobody else uses during this call
    IasCsrMatrixDev dA = {}, dB = {};
    IasDiaDev a_dia = {}, b_dia = {}, c_dia = {};
    IasEllDev a_ell = {}, b_ell = {};
    IasEll64Dev c_ell = {};
    IasCsr64Dev c_csr = {};
    int rc = IAS_OK;
    auto cleanup = [&]() {
        ias_free_dia_dev(&c_dia); ias_free_ell64_dev(&c_ell); ias_free_csr64_dev(&c_csr);
        ias_free_dia_dev(&a_dia); if (!alias) ias_free_dia_dev(&b_dia);
        ias_free_ell_dev(&a_ell); if (!alias) ias_free_ell_dev(&b_ell);
        ias_free_csr_dev(&dA); if (!alias) ias_free_csr_dev(&dB);
    };
#define AUTO_TRY(x) do { rc = (x); if (rc != IAS_OK) { cleanup(); return rc; } } while (0)

    cudaEventRecord(ev[0], s);
    int have_dA = 0;
    if (alias && !matnet && c.tune.e2e_pipeline != 0) {
        // banded A^2: upload, DIA multiply and download overlapped chunk by chunk (speculative, verified; see dia.cu)
        int done = 0;
        rc = auto_dia_pipelined(A, gate, out, &dA, &have_dA, &done);
        if (rc != IAS_OK) { if (have_dA) ias_free_csr_dev(&dA); return rc; }
        if (done) {
            clock_gettime(CLOCK_MONOTONIC, &t_out);
            out->ms_wall = (t_out.tv_sec - t_in.tv_sec) * 1e3 + (t_out.tv_nsec - t_in.tv_nsec) / 1e6;
            out->pipelined = 1;
            for (int k = 0; k < 6; ++k) out->ms_host[k] = out->ms_wall;
            return IAS_OK;
        }
        memset(out, 0, sizeof *out);
    }
    if (!have_dA) AUTO_TRY(ias_upload_csr(A, &dA));
    if (alias) dB = dA; else AUTO_TRY(ias_upload_csr(B, &dB));
    cudaEventRecord(ev[1], s);
    stamp(0);

    // ---- features (CPU/main.cpp:655-679).  The DIA conversion doubles as the diagonal census.
    double *f = out->features;
    AUTO_TRY(ias_getinfo1(&dA, f));
    if (alias) memcpy(f + 9, f, 9 * sizeof(double)); else AUTO_TRY(ias_getinfo1(&dB, f + 9));
    AUTO_TRY(ias_csr_to_dia(&dA, gate, &a_dia));
    if (alias) b_dia = a_dia; else AUTO_TRY(ias_csr_to_dia(&dB, gate, &b_dia));
    ias_getinfo2(dA.row, dA.col, a_dia.num_diagonals, f + 18);
    ias_getinfo2(dB.row, dB.col, b_dia.num_diagonals, f + 21);
    int wa = 0, wb = 0;
    AUTO_TRY(ias_max_row_nnz(&dA, &wa));
    if (alias) wb = wa; else AUTO_TRY(ias_max_row_nnz(&dB, &wb));
    ias_getinfo3(dA.row, dA.nnz, std::max(wa, 1), f + 24);
    ias_getinfo3(dB.row, dB.nnz, std::max(wb, 1), f + 25);
    const bool dia_ok = a_dia.choice && b_dia.choice;
    const bool ell_ok = ias_sizeof_ell(dA.row, wa) < gate * ias_sizeof_csr(dA.row, dA.nnz) &&
                        ias_sizeof_ell(dB.row, wb) < gate * ias_sizeof_csr(dB.row, dB.nnz);
    int cls = ias_select_format(f, dia_ok, ell_ok);
    if (matnet) {                                 // MatNet.Pred on the density images + features (CPU/MatNet.py:24-96)
        static long long img1[16384], img2[16384];
        AUTO_TRY(ias_density_image(&dA, img1));
        if (alias) memcpy(img2, img1, sizeof img1); else AUTO_TRY(ias_density_image(&dB, img2));
        int nf = 0, nc = 0, k = 0;
        double probs[8] = {0};
        ias_matnet_shape(matnet, &nf, &nc, nullptr);
        AUTO_TRY(ias_matnet_predict(matnet, img1, img2, f, &k, probs));
        if (nc == 5) {                            // 0 MKL 1 CSR 2 DIA 3 ELL 4 COO
            cls = (k == 0 || k == 4) ? 1 : k;
            if ((cls == 2 && !dia_ok) || (cls == 3 && !ell_ok)) cls = 1;
        } else cls = 1;                           // the GPU net names a library CSR SpGEMM
    }
    out->format = cls;
    out->row = dA.row; out->col = dB.col;
    cudaEventRecord(ev[2], s);
    stamp(1);

    // ---- conversion + multiply in the selected format, result into the pinned host arena
    void *base = nullptr;
    if (cls == 2) {
        cudaEventRecord(ev[3], s);                // DIA operands were built above (counted as selection + conversion)
        stamp(2);
        double ms = 0;
        AUTO_TRY(ias_dia_mul_dia_dev(&a_dia, &b_dia, &c_dia, &ms));
        const int nd = c_dia.num_diagonals;
        const size_t cells = (size_t)c_dia.row * nd;
        DBuf<double> rm;
        AUTO_TRY(rm.alloc(cells));
        if (cells) { dia_rows_major(c_dia.row, nd, c_dia.values_dev, rm.p, s); c.launches++; }
        cudaEventRecord(ev[4], s);
        stamp(3);
        const size_t span = (size_t)std::max(c_dia.row + c_dia.col - 1, 1);
        const size_t o_off = up256(cells * 8), o_ind = o_off + up256((size_t)std::max(nd, 1) * 4);
        AUTO_TRY(host_arena(o_ind + span * 4 + 256, &base));
        out->values = (double *)base;
        out->diagonal_offsets = (int *)((char *)base + o_off);
        out->diagonal_ind = (int *)((char *)base + o_ind);
        if (cells) cudaMemcpyAsync(out->values, rm.p, cells * 8, cudaMemcpyDeviceToHost, s);
        if (nd) cudaMemcpyAsync(out->diagonal_offsets, c_dia.diagonal_offsets_dev, (size_t)nd * 4, cudaMemcpyDeviceToHost, s);
        if (c_dia.row + c_dia.col - 1 > 0) cudaMemcpyAsync(out->diagonal_ind, c_dia.diagonal_ind_dev, span * 4, cudaMemcpyDeviceToHost, s);
        cudaEventRecord(ev[5], s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); return fail_cuda(cudaGetLastError(), "download DIA", __FILE__, __LINE__); }
        out->num_diagonals = nd;
        out->nnz = (long long)cells;
        out->d2h_bytes = (long long)(cells * 8 + (size_t)nd * 4 + span * 4);
    } else if (cls == 3) {
        AUTO_TRY(ias_csr_to_ell(&dA, gate, &a_ell));
        if (alias) b_ell = a_ell; else AUTO_TRY(ias_csr_to_ell(&dB, gate, &b_ell));
        cudaEventRecord(ev[3], s);
        double ms = 0;
        AUTO_TRY(ias_ell_mul_ell_dev64(&a_ell, &b_ell, &c_ell, &ms));
        cudaEventRecord(ev[4], s);
        const size_t cells = (size_t)c_ell.row * c_ell.max_nnz_per_row;
        const size_t o_ci = up256(cells * 8), o_nr = o_ci + up256(cells * 4);
        AUTO_TRY(host_arena(o_nr + (size_t)std::max(c_ell.row, 1) * 4 + 256, &base));
        out->values = (double *)base;
        out->col_ind = (int *)((char *)base + o_ci);
        out->nnz_row = (int *)((char *)base + o_nr);
        if (cells) {
            cudaMemcpyAsync(out->values, c_ell.values_dev, cells * 8, cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(out->col_ind, c_ell.col_ind_dev, cells * 4, cudaMemcpyDeviceToHost, s);
        }
        if (c_ell.row) cudaMemcpyAsync(out->nnz_row, c_ell.nnz_row_dev, (size_t)c_ell.row * 4, cudaMemcpyDeviceToHost, s);
        cudaEventRecord(ev[5], s);
        if (cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); return fail_cuda(cudaGetLastError(), "download ELL", __FILE__, __LINE__); }
        out->max_nnz_per_row = c_ell.max_nnz_per_row;
        out->nnz = c_ell.nnz;
        out->d2h_bytes = (long long)(cells * 12 + (size_t)c_ell.row * 4);
    } else {
        cudaEventRecord(ev[3], s);
        IasSpgemmStats st;
        AUTO_TRY(ias_csr_mul_csr_dev64(&dA, &dB, &c_csr, &st));
        cudaEventRecord(ev[4], s);
        const size_t b_rp = 8 * ((size_t)c_csr.row + 1), b_v = 8 * (size_t)c_csr.nnz, b_ci = 4 * (size_t)c_csr.nnz;
        const size_t o_v = up256(b_rp), o_ci = o_v + up256(b_v);
        AUTO_TRY(host_arena(o_ci + b_ci + 256, &base));
        out->row_ptr = (long long *)base;
        out->values = (double *)((char *)base + o_v);
        out->col_ind = (int *)((char *)base + o_ci);
        AUTO_TRY(ias_download_csr64(&c_csr, out->row_ptr, out->col_ind, out->values));
        cudaEventRecord(ev[5], s);
        cudaStreamSynchronize(s);
        out->nnz = c_csr.nnz;
        out->d2h_bytes = (long long)(b_rp + b_v + b_ci);
    }
    out->h2d_bytes = (long long)(4 * ((size_t)A->row + 1) + 12 * (size_t)A->nnz) +
                     (alias ? 0 : (long long)(4 * ((size_t)B->row + 1) + 12 * (size_t)B->nnz));
    out->ms_h2d = ms_between(ev[0], ev[1]);
    out->ms_select = ms_between(ev[1], ev[2]);
    out->ms_convert = ms_between(ev[2], ev[3]);
    out->ms_multiply = ms_between(ev[3], ev[4]);
    out->ms_d2h = ms_between(ev[4], ev[5]);
    stamp(4);
    cleanup();
    stamp(5);
    clock_gettime(CLOCK_MONOTONIC, &t_out);
    out->ms_wall = (t_out.tv_sec - t_in.tv_sec) * 1e3 + (t_out.tv_nsec - t_in.tv_nsec) / 1e6;
#undef AUTO_TRY
    return IAS_OK;
}

}  // extern "C"
