// dia.cu -- DIA format: converter and the DIA x DIA kernel (Algorithm 3).
//
// Replaces CSRtoDIA (CPU/detail/dia/common_dia.h:29-96), DIA_mul_DIA (dia:101-195) and the
// never-called DIA_MUL_DIA_DEV chain (GPU/detail/dia_dev/common_dia_dev.h:27-182: phase1 flags
// diagonals with a row loop, DIA_sum<<<1,1>>> compacts serially, phase3 walks row-major values with
// a stride of num_diagonals).  Here:
//   - values are DIAGONAL-MAJOR on device (values[slot*rows + i]) so a warp reads/writes 256
//     contiguous bytes per diagonal;
//   - the set of output diagonals is a function of the offsets and the matrix bounds alone
//     (dia:104-140 never looks at values), so it is computed on the host in O(dA*dB);
//   - one pass writes every output diagonal once, accumulating all contributing (a,b) pairs in a
//     register: HBM traffic = 8*n*(dA + dB + dC) bytes, the algorithmic minimum.
#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"

using namespace ias;

namespace {

__global__ void __launch_bounds__(256) k_flag_diagonals(int rows, const int *__restrict__ rp, const int *__restrict__ ci,
                                                        int *__restrict__ flags /* rows+cols, index (rows-i)+j */)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int pe = rp[i + 1];
    for (int p = rp[i]; p < pe; ++p) {
        int m = (rows - i) + ci[p];
        if (!flags[m]) flags[m] = 1;
    }
}

// slot_of = exclusive scan of flags.  offsets[slot] = m - rows; diagonal_ind[m-1] = slot (0 when absent)
__global__ void __launch_bounds__(256) k_number_diagonals(int span, int rows, const int *__restrict__ flags,
                                                          const int *__restrict__ slot_of, int *__restrict__ offsets,
                                                          int *__restrict__ diag_ind)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= span) return;
    int present = flags[m];
    if (present) offsets[slot_of[m]] = m - rows;
    if (m >= 1) diag_ind[m - 1] = present ? slot_of[m] : 0;
}

__global__ void __launch_bounds__(256) k_fill_dia(int rows, const int *__restrict__ rp, const int *__restrict__ ci,
                                                  const double *__restrict__ v, const int *__restrict__ slot_of,
                                                  double *__restrict__ values /* diagonal-major */)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int pe = rp[i + 1];
    for (int p = rp[i]; p < pe; ++p) {           // in order: a duplicate (i,j) overwrites, as dia:75-90 does
        int slot = slot_of[(rows - i) + ci[p]];
        values[(size_t)slot * rows + i] = v[p];
    }
}

// C[d][i] = sum over pairs (a,b) of diagonal d:  A[a][i] * B[b][i + offA[a]]   (dia:162-193)
// ncu on the first version (per-pair offset lookups, four 64-bit bound checks): 1635 instructions per row,
// issue slots 81 % busy -- instruction bound at 48 % of the HBM roofline.  Now every pair carries its two
// base indices and the row interval on which it contributes (all computed on the host), and a thread owns
// two rows, so a pair costs one 24-byte shared-memory read, two compares and two loads + one FMA per row.
struct DiaPair {
    long long a0;      // a * rows                : A value index = a0 + i
    long long b0;      // b * rows(B) + offA[a]   : B value index = b0 + i
    int lo, hi;        // the pair contributes on rows lo <= i < hi
    int pad0, pad1;
};

template <int BLOCK, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(BLOCK) k_dia_mul_dia(int row0, int rows /* one past the last row */, int c_nd, const double *__restrict__ a_val,
                                                       const double *__restrict__ b_val,
                                                       const int *__restrict__ pair_start /* c_nd+1 */,
                                                       const DiaPair *__restrict__ pairs, int npairs,
                                                       double *__restrict__ c_val)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const DiaPair *s_pairs = pairs;
    const int *s_start = pair_start;
    if (TABLES_IN_SMEM) {                      // the usual case: a few diagonals, tables of a few hundred bytes
        DiaPair *w_pairs = reinterpret_cast<DiaPair *>(sm_raw);
        int *w_start = reinterpret_cast<int *>(w_pairs + npairs);
        for (int t = threadIdx.x; t < npairs; t += BLOCK) w_pairs[t] = pairs[t];
        for (int t = threadIdx.x; t <= c_nd; t += BLOCK) w_start[t] = pair_start[t];
        __syncthreads();
        s_pairs = w_pairs; s_start = w_start;
    }
    const size_t out_rows = (size_t)(rows - row0);          // C holds rows [row0, rows): values[d * out_rows + (i - row0)]
    for (long long base = row0 + (long long)blockIdx.x * (2 * BLOCK); base < rows; base += (long long)gridDim.x * (2 * BLOCK)) {
        const long long i0 = base + threadIdx.x, i1 = i0 + BLOCK;
        const bool v0 = i0 < rows, v1 = i1 < rows;
        for (int d = 0; d < c_nd; ++d) {
            double acc0 = 0.0, acc1 = 0.0;
            const int pe = s_start[d + 1];
            for (int p = s_start[d]; p < pe; ++p) {
                const DiaPair P = s_pairs[p];
                if (i0 >= P.lo && i0 < P.hi) acc0 += __ldg(a_val + P.a0 + i0) * __ldg(b_val + P.b0 + i0);
                if (i1 >= P.lo && i1 < P.hi) acc1 += __ldg(a_val + P.a0 + i1) * __ldg(b_val + P.b0 + i1);
            }
            if (v0) c_val[(size_t)d * out_rows + (i0 - row0)] = acc0;
            if (v1) c_val[(size_t)d * out_rows + (i1 - row0)] = acc1;
        }
    }
}

// Same product with 128-bit accesses: a thread owns two ADJACENT rows (i, i + 1), i even.  ncu on the scalar kernel:
// issue slots 76 % busy, DRAM at 0.76 of the copy peak -- instruction bound.  Here a pair whose A and B indices are
// even for even i (flag bit 0 / bit 1 of DiaPair::pad0, set on the host) costs one LDG.128 per operand and two FMAs
// for two rows; odd offsets (the +-1 diagonals of a stencil) fall back to two 64-bit loads for that operand.
// Requires an even row0 and an even number of output rows per diagonal (so that every C store is 16-byte aligned).
template <int BLOCK, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(BLOCK) k_dia_mul_dia_v2(int row0, int rows /* one past the last row */, int c_nd, const double *__restrict__ a_val,
                                                          const double *__restrict__ b_val,
                                                          const int *__restrict__ pair_start /* c_nd+1 */,
                                                          const DiaPair *__restrict__ pairs, int npairs,
                                                          double *__restrict__ c_val)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const DiaPair *s_pairs = pairs;
    const int *s_start = pair_start;
    if (TABLES_IN_SMEM) {
        DiaPair *w_pairs = reinterpret_cast<DiaPair *>(sm_raw);
        int *w_start = reinterpret_cast<int *>(w_pairs + npairs);
        for (int t = threadIdx.x; t < npairs; t += BLOCK) w_pairs[t] = pairs[t];
        for (int t = threadIdx.x; t <= c_nd; t += BLOCK) w_start[t] = pair_start[t];
        __syncthreads();
        s_pairs = w_pairs; s_start = w_start;
    }
    const size_t out_rows = (size_t)(rows - row0);
    for (long long base = row0 + (long long)blockIdx.x * (2 * BLOCK); base < rows; base += (long long)gridDim.x * (2 * BLOCK)) {
        const long long i = base + 2 * threadIdx.x;                    // rows i and i + 1 (rows is even: both or neither exist)
        if (i >= rows) continue;
        for (int d = 0; d < c_nd; ++d) {
            double acc0 = 0.0, acc1 = 0.0;
            const int pe = s_start[d + 1];
            for (int p = s_start[d]; p < pe; ++p) {
                const DiaPair P = s_pairs[p];
                if (i >= P.lo && i + 1 < P.hi) {                       // both rows contribute: the common case
                    double a0, a1, b0, b1;
                    if (P.pad0 & 1) { const double2 t = __ldg(reinterpret_cast<const double2 *>(a_val + P.a0 + i)); a0 = t.x; a1 = t.y; }
                    else { a0 = __ldg(a_val + P.a0 + i); a1 = __ldg(a_val + P.a0 + i + 1); }
                    if (P.pad0 & 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(b_val + P.b0 + i)); b0 = t.x; b1 = t.y; }
                    else { b0 = __ldg(b_val + P.b0 + i); b1 = __ldg(b_val + P.b0 + i + 1); }
                    acc0 += a0 * b0;
                    acc1 += a1 * b1;
                } else {
                    if (i >= P.lo && i < P.hi) acc0 += __ldg(a_val + P.a0 + i) * __ldg(b_val + P.b0 + i);
                    if (i + 1 >= P.lo && i + 1 < P.hi) acc1 += __ldg(a_val + P.a0 + i + 1) * __ldg(b_val + P.b0 + i + 1);
                }
            }
            *reinterpret_cast<double2 *>(c_val + (size_t)d * out_rows + (i - row0)) = make_double2(acc0, acc1);
        }
    }
}

// diagonal_ind[offset + rows - 1] = slot for the c_nd present diagonals (the rest stays 0)
__global__ void k_scatter_diag_ind(int c_nd, int rows, const int *__restrict__ offsets, int *__restrict__ diag_ind)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < c_nd) diag_ind[(long long)offsets[d] + rows - 1] = d;
}

// diagonal-major -> the reference's row-major [i*nd + slot] (for downloads / parity checks)
__global__ void __launch_bounds__(256) k_dia_to_row_major(int rows, int nd, const double *__restrict__ in, double *__restrict__ out)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)rows * nd;
    if (t >= n) return;
    size_t i = t / nd, s = t % nd;
    out[t] = in[s * rows + i];
}

// the reference's row-major [i*nd + slot] -> diagonal-major
__global__ void __launch_bounds__(256) k_dia_to_diag_major(int rows, int nd, const double *__restrict__ in, double *__restrict__ out)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)rows * nd;
    if (t >= n) return;
    size_t s = t / rows, i = t % rows;
    out[t] = in[i * nd + s];
}

int diag_census(const IasCsrMatrixDev *A, DBuf<int> &flags, DBuf<int> &slot_of, int *nd)
{
    Ctx &c = ctx();
    int span = A->row + A->col;                     // map index (rows - i) + j lies in [1, rows+cols-1]
    IAS_TRY(flags.alloc((size_t)span + 1));
    IAS_TRY(slot_of.alloc((size_t)span + 1));
    IAS_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(int) * ((size_t)span + 1), c.stream));
    if (A->row) IAS_LAUNCH(k_flag_diagonals, grid_for(A->row, 256), 256, 0, A->row, A->row_ind_dev, A->col_ind_dev, flags.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags.p, slot_of.p, span + 1, c.stream));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags.p, slot_of.p, span + 1, c.stream));
    c.launches += 2;
    IAS_CUDA(cudaMemcpyAsync(nd, slot_of.p + span, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

}  // namespace

extern "C" {

int ias_count_diagonals(const IasCsrMatrixDev *A, int *num_diagonals)
{
    IAS_TRY(ensure_init());
    if (!A || !num_diagonals) return fail(IAS_E_ARG, "NULL");
    DBuf<int> flags, slot_of;
    return diag_census(A, flags, slot_of, num_diagonals);
}

int ias_csr_to_dia(const IasCsrMatrixDev *A, double gate, IasDiaDev *out)
{
    IAS_TRY(ensure_init());
    if (!A || !out) return fail(IAS_E_ARG, "NULL");
    Ctx &c = ctx();
    memset(out, 0, sizeof *out);
    out->row = A->row; out->col = A->col;
    DBuf<int> flags, slot_of;
    int nd = 0;
    IAS_TRY(diag_census(A, flags, slot_of, &nd));
    out->num_diagonals = nd;
    // size gate of the reference (dia:56 uses 50x on the CPU, GPU/detail/dia/common_dia.h:51 uses 20x)
    if (!(ias_sizeof_dia(A->row, A->col, nd) < gate * ias_sizeof_csr(A->row, A->nnz))) {
        out->choice = false;
        return IAS_OK;
    }
    if (nd > 65535) { out->choice = false; return IAS_OK; }   // pair tables index diagonals with 16 bits
    out->choice = true;
    int span = A->row + A->col;
    DBuf<int> di, off;
    DBuf<double> val;
    IAS_TRY(di.alloc((size_t)std::max(span - 1, 1)));
    IAS_TRY(off.alloc((size_t)std::max(nd, 1)));
    IAS_TRY(val.alloc((size_t)A->row * nd));
    IAS_CUDA(cudaMemsetAsync(val.p, 0, sizeof(double) * (size_t)A->row * nd, c.stream));
    IAS_LAUNCH(k_number_diagonals, grid_for(span, 256), 256, 0, span, A->row, flags.p, slot_of.p, off.p, di.p);
    if (A->row) IAS_LAUNCH(k_fill_dia, grid_for(A->row, 256), 256, 0, A->row, A->row_ind_dev, A->col_ind_dev, A->values_dev, slot_of.p, val.p);
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    out->diagonal_ind_dev = di.release(); out->diagonal_offsets_dev = off.release(); out->values_dev = val.release();
    return IAS_OK;
}

int ias_free_dia_dev(IasDiaDev *m)
{
    if (!m) return IAS_OK;
    dfree(m->diagonal_ind_dev); dfree(m->diagonal_offsets_dev); dfree(m->values_dev);
    m->diagonal_ind_dev = nullptr; m->diagonal_offsets_dev = nullptr; m->values_dev = nullptr;
    return IAS_OK;
}

int ias_dia_mul_dia_dev(const IasDiaDev *A, const IasDiaDev *B, IasDiaDev *C, double *elapsed_ms)
{
    if (!A) return fail(IAS_E_ARG, "NULL");
    return ias_dia_mul_dia_rows_dev(A, B, 0, A->row, C, elapsed_ms);
}

int ias_dia_mul_dia_rows_dev(const IasDiaDev *A, const IasDiaDev *B, int r0, int r1, IasDiaDev *C, double *elapsed_ms)
{
    IAS_TRY(ensure_init());
    if (!A || !B || !C) return fail(IAS_E_ARG, "NULL");
    if (r0 < 0 || r1 > A->row || r0 > r1) return fail(IAS_E_ARG, "row range [%d,%d) outside A (%d rows)", r0, r1, A->row);
    if (!A->choice || !B->choice) return fail(IAS_E_GATE, "DIA operand was rejected by the size gate (choice == false)");
    if (A->col > B->row) return fail(IAS_E_ARG, "shape mismatch: A is %dx%d, B is %dx%d", A->row, A->col, B->row, B->col);
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    memset(C, 0, sizeof *C);
    const int nrows = r1 - r0;
    C->row = nrows; C->col = B->col; C->choice = true;
    IAS_CUDA(cudaEventRecord(c.ev[0], s));

    // offsets are tiny: fetch them and enumerate the reachable output diagonals on the host.
    // (a,b) reaches diagonal oA+oB iff some row i has 0 <= i+oA < a_cols and 0 <= i+oA+oB < b_cols (dia:110-131)
    std::vector<int> ao(A->num_diagonals), bo(B->num_diagonals);
    if (!ao.empty()) IAS_CUDA(cudaMemcpyAsync(ao.data(), A->diagonal_offsets_dev, sizeof(int) * ao.size(), cudaMemcpyDeviceToHost, s));
    if (!bo.empty()) IAS_CUDA(cudaMemcpyAsync(bo.data(), B->diagonal_offsets_dev, sizeof(int) * bo.size(), cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    struct Pair { long long off; DiaPair dp; };
    std::vector<Pair> pairs;
    pairs.reserve(ao.size() * bo.size());
    for (size_t a = 0; a < ao.size(); ++a)
        for (size_t b = 0; b < bo.size(); ++b) {
            long long oa = ao[a], ob = bo[b];
            long long lo = std::max<long long>(0, std::max(-oa, -oa - ob));
            long long hi = std::min<long long>(A->row, std::min<long long>((long long)A->col - oa, (long long)B->col - oa - ob));
            if (lo < hi)
            {
                // bit 0 / bit 1: the A / B value index of an even row is even (128-bit loads in k_dia_mul_dia_v2)
                const long long a0 = (long long)a * A->row, b0 = (long long)b * B->row + oa;
                const int al = ((a0 & 1) == 0 ? 1 : 0) | ((b0 & 1) == 0 ? 2 : 0);
                pairs.push_back({oa + ob, DiaPair{a0, b0, (int)lo, (int)hi, al, 0}});
            }
        }
    std::stable_sort(pairs.begin(), pairs.end(), [](const Pair &x, const Pair &y) { return x.off < y.off; });
    std::vector<int> c_off, pstart;
    std::vector<DiaPair> pab(pairs.size());
    for (size_t p = 0; p < pairs.size(); ++p) {
        if (p == 0 || pairs[p].off != pairs[p - 1].off) { c_off.push_back((int)pairs[p].off); pstart.push_back((int)p); }
        pab[p] = pairs[p].dp;
    }
    pstart.push_back((int)pairs.size());
    int c_nd = (int)c_off.size();
    C->num_diagonals = c_nd;

    // (diagonal_ind keeps the index space of the whole product, rows(A) + cols(B) - 1 entries, for every row block)
    int span = A->row + B->col - 1;
    DBuf<int> di, off, d_pstart;
    DBuf<DiaPair> d_pairs;
    DBuf<double> val;
    // diagonal_ind spans rows(A) + cols(B) - 1 entries whatever the row block: a block of a larger product (multi-GPU)
    // does not carry it -- zeroing a gigabyte per multiply is what cost the 8-GPU weak-scaling run 25 %
    const bool full = r0 == 0 && r1 == A->row;
    if (full) IAS_TRY(di.alloc((size_t)std::max(span, 1)));
    IAS_TRY(off.alloc((size_t)std::max(c_nd, 1)));
    IAS_TRY(d_pstart.alloc(pstart.size()));
    IAS_TRY(d_pairs.alloc(std::max<size_t>(pab.size(), 1)));
    IAS_TRY(val.alloc((size_t)nrows * c_nd));
    if (full) IAS_CUDA(cudaMemsetAsync(di.p, 0, sizeof(int) * (size_t)std::max(span, 1), s));
    if (c_nd) {
        IAS_CUDA(cudaMemcpyAsync(off.p, c_off.data(), sizeof(int) * c_nd, cudaMemcpyHostToDevice, s));
        if (full) IAS_LAUNCH(k_scatter_diag_ind, grid_for(c_nd, 256), 256, 0, c_nd, A->row, off.p, di.p);     // dia:150-158
    }
    IAS_CUDA(cudaMemcpyAsync(d_pstart.p, pstart.data(), sizeof(int) * pstart.size(), cudaMemcpyHostToDevice, s));
    if (!pab.empty()) IAS_CUDA(cudaMemcpyAsync(d_pairs.p, pab.data(), sizeof(DiaPair) * pab.size(), cudaMemcpyHostToDevice, s));

    size_t sm = sizeof(DiaPair) * pab.size() + sizeof(int) * ((size_t)c_nd + 1);
    // 128-bit variant: even first row, even row count (C stores), 16-byte aligned value arrays
    const bool vec = c.tune.dia_vec != 0 && (r0 & 1) == 0 && (nrows & 1) == 0 && sm <= 32 * 1024 &&
                     ((uintptr_t)A->values_dev & 15) == 0 && ((uintptr_t)B->values_dev & 15) == 0;
    if (nrows && c_nd && vec) {
        constexpr int BLOCK = 256;
        unsigned grid = (unsigned)std::min<long long>(grid_for(nrows, 2 * BLOCK), (long long)c.sm_count * 8 * 64);
        IAS_LAUNCH((k_dia_mul_dia_v2<BLOCK, true>), grid, BLOCK, sm, r0, r1, c_nd, A->values_dev, B->values_dev, d_pstart.p, d_pairs.p,
                   (int)pab.size(), val.p);
    } else if (nrows && c_nd) {
        constexpr int BLOCK = 256;
        unsigned grid = (unsigned)std::min<long long>(grid_for(nrows, 2 * BLOCK), (long long)c.sm_count * 8 * 64);
        if (sm <= 32 * 1024) {
            IAS_LAUNCH((k_dia_mul_dia<BLOCK, true>), grid, BLOCK, sm, r0, r1, c_nd, A->values_dev, B->values_dev, d_pstart.p, d_pairs.p,
                       (int)pab.size(), val.p);
        } else {                               // many diagonals: the tables stay in global memory (L1/L2 resident)
            IAS_LAUNCH((k_dia_mul_dia<BLOCK, false>), grid, BLOCK, 0, r0, r1, c_nd, A->values_dev, B->values_dev, d_pstart.p, d_pairs.p,
                       (int)pab.size(), val.p);
        }
    }
    IAS_CUDA(cudaEventRecord(c.ev[1], s));
    IAS_CUDA(cudaStreamSynchronize(s));
    if (elapsed_ms) { float ms = 0; cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]); *elapsed_ms = ms; }
    C->diagonal_ind_dev = di.release(); C->diagonal_offsets_dev = off.release(); C->values_dev = val.release();
    return IAS_OK;
}

int ias_dia_relayout(const IasDiaDev *in, int to_row_major, IasDiaDev *out)
{
    IAS_TRY(ensure_init());
    if (!in || !out) return fail(IAS_E_ARG, "NULL");
    if (!in->choice) return fail(IAS_E_GATE, "DIA matrix was rejected by the size gate");
    Ctx &c = ctx();
    IasDiaDev r = *in;
    r.diagonal_ind_dev = nullptr; r.diagonal_offsets_dev = nullptr; r.values_dev = nullptr;
    int span = std::max(in->row + in->col - 1, 1);
    size_t n = (size_t)in->row * in->num_diagonals;
    DBuf<int> di, off;
    DBuf<double> val;
    IAS_TRY(di.alloc((size_t)span));
    IAS_TRY(off.alloc((size_t)std::max(in->num_diagonals, 1)));
    IAS_TRY(val.alloc(n));
    if (in->row + in->col - 1 > 0 && in->diagonal_ind_dev) IAS_CUDA(cudaMemcpyAsync(di.p, in->diagonal_ind_dev, sizeof(int) * (size_t)(in->row + in->col - 1), cudaMemcpyDeviceToDevice, c.stream));
    if (in->num_diagonals) IAS_CUDA(cudaMemcpyAsync(off.p, in->diagonal_offsets_dev, sizeof(int) * (size_t)in->num_diagonals, cudaMemcpyDeviceToDevice, c.stream));
    if (n) {
        if (to_row_major) IAS_LAUNCH(k_dia_to_row_major, grid_for((long long)n, 256), 256, 0, in->row, in->num_diagonals, in->values_dev, val.p);
        else IAS_LAUNCH(k_dia_to_diag_major, grid_for((long long)n, 256), 256, 0, in->row, in->num_diagonals, in->values_dev, val.p);
    }
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    r.diagonal_ind_dev = di.release(); r.diagonal_offsets_dev = off.release(); r.values_dev = val.release();
    *out = r;
    return IAS_OK;
}

int ias_download_dia(const IasDiaDev *d, int *diagonal_ind, int *diagonal_offsets, double *values_row_major)
{
    IAS_TRY(ensure_init());
    if (!d) return fail(IAS_E_ARG, "NULL");
    if (!d->choice) return fail(IAS_E_GATE, "DIA matrix was rejected by the size gate");
    Ctx &c = ctx();
    int span = d->row + d->col - 1;
    if (diagonal_ind && span > 0 && d->diagonal_ind_dev) IAS_CUDA(cudaMemcpyAsync(diagonal_ind, d->diagonal_ind_dev, sizeof(int) * span, cudaMemcpyDeviceToHost, c.stream));
    if (diagonal_offsets && d->num_diagonals) IAS_CUDA(cudaMemcpyAsync(diagonal_offsets, d->diagonal_offsets_dev, sizeof(int) * d->num_diagonals, cudaMemcpyDeviceToHost, c.stream));
    size_t n = (size_t)d->row * d->num_diagonals;
    if (values_row_major && n) {
        DBuf<double> rm;
        IAS_TRY(rm.alloc(n));
        IAS_LAUNCH(k_dia_to_row_major, grid_for((long long)n, 256), 256, 0, d->row, d->num_diagonals, d->values_dev, rm.p);
        IAS_CUDA(cudaMemcpyAsync(values_row_major, rm.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
        IAS_CUDA(cudaStreamSynchronize(c.stream));
    }
    IAS_CUDA(cudaStreamSynchronize(c.stream));
    return IAS_OK;
}

}  // extern "C"
