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

// ---- kernels of the pipelined host path (row ranges of an operand that is still being uploaded)
__global__ void __launch_bounds__(256) k_flag_diagonals_rows(int r0, int r1, int rows, const int *__restrict__ rp, const int *__restrict__ ci,
                                                             int *__restrict__ flags)
{
    int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r1) return;
    int pe = rp[i + 1];
    for (int p = rp[i]; p < pe; ++p) {
        int m = (rows - i) + ci[p];
        if (!flags[m]) flags[m] = 1;
    }
}
// fill rows [r0, r1) of a DIA operand whose diagonal set was taken from the first chunk: an entry on another diagonal
// raises *bad (the speculation failed; the caller falls back to the unpipelined path)
__global__ void __launch_bounds__(256) k_fill_dia_rows(int r0, int r1, int rows, const int *__restrict__ rp, const int *__restrict__ ci,
                                                       const double *__restrict__ v, const int *__restrict__ spec_flags,
                                                       const int *__restrict__ slot_of, double *__restrict__ values, int *__restrict__ bad)
{
    int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r1) return;
    int pe = rp[i + 1];
    for (int p = rp[i]; p < pe; ++p) {
        int m = (rows - i) + ci[p];
        if (!spec_flags[m]) { *bad = 1; continue; }
        values[(size_t)slot_of[m] * rows + i] = v[p];
    }
}
__global__ void __launch_bounds__(256) k_flags_differ(int n, const int *__restrict__ a, const int *__restrict__ b, int *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (a[i] != 0) != (b[i] != 0)) out[1] = 1;
}
// tiled transpose diagonal-major -> row-major (both sides coalesced)
__global__ void __launch_bounds__(256) k_dia_rows_major(int rows, int nd, const double *__restrict__ in, double *__restrict__ out)
{
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 8 rows of 32 threads
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, i = i0 + tx;
        if (s < nd && i < rows) tile[r][tx] = in[(size_t)s * rows + i];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, s = s0 + tx;
        if (i < rows && s < nd) out[(size_t)i * nd + s] = tile[tx][r];
    }
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

namespace ias {

void dia_rows_major(int rows, int nd, const double *in, double *out, cudaStream_t s)
{
    if (rows <= 0 || nd <= 0) return;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((nd + 31) / 32));
    k_dia_rows_major<<<grid, 256, 0, s>>>(rows, nd, in, out);
}

// A^2 of a banded host operand with upload, multiply and download overlapped (PCIe is full duplex):
//   the CSR arrays go up in row chunks on one stream; the diagonal set is taken from the FIRST chunk (speculation) so that
//   chunk k can be converted to DIA and chunk k-1 multiplied (ias_dia_mul_dia_rows_dev: it needs the B rows within the
//   band around its own rows, i.e. one chunk ahead) while later chunks are still in flight; each finished row block of C
//   is transposed to the reference's row-major layout and goes down on a third stream into the pinned result.
//   Afterwards the speculation is VERIFIED: the census of all rows must equal the first chunk's, no entry may have
//   fallen outside it, and the selector run on the complete features must still say DIA -- otherwise *done stays 0 and
//   the caller runs the plain path on the operand that is on the device by then (*have_dA).
// A^2 needs all of B before the first row of C only when B's rows are referenced at random; a band references its
// neighbourhood, and that is what makes the overlap legal here.
int auto_dia_pipelined(const IasCsrMatrix *A, double gate, IasAutoResult *out, IasCsrMatrixDev *dA, int *have_dA, int *done)
{
    *done = 0; *have_dA = 0;
    Ctx &c = ctx();
    const int rows = A->row, cols = A->col;
    const long long nnz = A->nnz;
    if (rows != cols || rows < (1 << 16) || nnz <= 0) return IAS_OK;
    IAS_TRY(ensure_pipe_streams());
    cudaStream_t s = c.stream, s_up = c.s_up, s_down = c.s_down;
    cudaEvent_t *ev_up = c.ev_pipe, *ev_c = c.ev_pipe + 32, *ev_d = c.ev_pipe + 64;
    const int K = 16;
    int chunk = ((rows + K - 1) / K + 1) & ~1;                      // even: keeps every block 16-byte aligned
    const int nchunk = (rows + chunk - 1) / chunk;

    // ---- device CSR arrays; uploads in row chunks on s_up
    DBuf<int> rp, ci;
    DBuf<double> v;
    IAS_TRY(rp.alloc((size_t)rows + 1));
    IAS_TRY(ci.alloc((size_t)nnz));
    IAS_TRY(v.alloc((size_t)nnz));
    const int span = rows + cols;
    DBuf<int> flags_spec, flags_full, slot_of, bad, a_off;
    IAS_TRY(flags_spec.alloc((size_t)span + 1));
    IAS_TRY(flags_full.alloc((size_t)span + 1));
    IAS_TRY(slot_of.alloc((size_t)span + 1));
    IAS_TRY(bad.alloc(2));
    IAS_CUDA(cudaMemsetAsync(flags_spec.p, 0, sizeof(int) * ((size_t)span + 1), s));
    IAS_CUDA(cudaMemsetAsync(flags_full.p, 0, sizeof(int) * ((size_t)span + 1), s));
    IAS_CUDA(cudaMemsetAsync(bad.p, 0, 2 * sizeof(int), s));
    IAS_CUDA(cudaStreamSynchronize(s));                              // the buffers exist before another stream writes them
    IAS_CUDA(cudaMemcpyAsync(rp.p, A->row_ind, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice, s_up));
    for (int k = 0; k < nchunk; ++k) {
        const int r0 = k * chunk, r1 = std::min(rows, r0 + chunk);
        const size_t e0 = (size_t)A->row_ind[r0], e1 = (size_t)A->row_ind[r1];
        if (e1 > e0) {
            IAS_CUDA(cudaMemcpyAsync(ci.p + e0, A->col_ind + e0, sizeof(int) * (e1 - e0), cudaMemcpyHostToDevice, s_up));
            IAS_CUDA(cudaMemcpyAsync(v.p + e0, A->values + e0, sizeof(double) * (e1 - e0), cudaMemcpyHostToDevice, s_up));
        }
        IAS_CUDA(cudaEventRecord(ev_up[k], s_up));
    }
    dA->choice = true; dA->row = rows; dA->col = cols; dA->nnz = (int)nnz;
    dA->row_ind_dev = rp.p; dA->col_ind_dev = ci.p; dA->values_dev = v.p;
    auto give_up = [&]() {                                            // hand the (completely uploaded) operand to the caller
        cudaStreamSynchronize(s_up); cudaStreamSynchronize(s_down); cudaStreamSynchronize(s);
        rp.release(); ci.release(); v.release();
        *have_dA = 1;
        return IAS_OK;
    };

    // ---- speculation: the diagonals of the first chunk
    IAS_CUDA(cudaStreamWaitEvent(s, ev_up[0], 0));
    const int c0_rows = std::min(rows, chunk);
    IAS_LAUNCH(k_flag_diagonals_rows, grid_for(c0_rows, 256), 256, 0, 0, c0_rows, rows, rp.p, ci.p, flags_spec.p);
    size_t tb = 0;
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags_spec.p, slot_of.p, span + 1, s));
    DBuf<char> tmp;
    IAS_TRY(tmp.alloc(tb));
    IAS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags_spec.p, slot_of.p, span + 1, s));
    c.launches += 2;
    int nd = 0;
    IAS_CUDA(cudaMemcpyAsync(&nd, slot_of.p + span, sizeof(int), cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    const bool plausible = nd > 0 && nd <= 4096 && ias_sizeof_dia(rows, cols, nd) < gate * ias_sizeof_csr(rows, nnz) &&
                           (double)nnz / ((double)nd * rows) > 0.5;
    if (!plausible) return give_up();
    IasDiaDev a_dia;
    memset(&a_dia, 0, sizeof a_dia);
    a_dia.choice = true; a_dia.row = rows; a_dia.col = cols; a_dia.num_diagonals = nd;
    DBuf<int> a_di;
    DBuf<double> a_val;
    IAS_TRY(a_off.alloc((size_t)nd));
    IAS_TRY(a_di.alloc((size_t)std::max(span - 1, 1)));
    IAS_TRY(a_val.alloc((size_t)rows * nd));
    IAS_CUDA(cudaMemsetAsync(a_val.p, 0, sizeof(double) * (size_t)rows * nd, s));
    IAS_LAUNCH(k_number_diagonals, grid_for(span, 256), 256, 0, span, rows, flags_spec.p, slot_of.p, a_off.p, a_di.p);
    a_dia.diagonal_ind_dev = a_di.p; a_dia.diagonal_offsets_dev = a_off.p; a_dia.values_dev = a_val.p;
    {
        // block j is multiplied when block j+1 has been converted: the band must not reach further than one chunk
        std::vector<int> h_aoff((size_t)nd);
        IAS_CUDA(cudaMemcpyAsync(h_aoff.data(), a_off.p, sizeof(int) * (size_t)nd, cudaMemcpyDeviceToHost, s));
        IAS_CUDA(cudaStreamSynchronize(s));
        long long reach = 0;
        for (int o : h_aoff) reach = std::max<long long>(reach, o);
        if (2 * reach >= chunk) return give_up();                   // (B's rows i + oA, and C = A*B reaches oA + oB)
    }

    // ---- chunks: convert k, multiply k-1, send k-1 down
    DBuf<double> stage[2];
    int ndc = 0;
    void *base = nullptr;
    double *h_vals = nullptr;
    int *h_off = nullptr, *h_ind = nullptr;
    size_t o_off = 0, o_ind = 0;
    auto multiply_block = [&](int j) -> int {
        const int r0 = j * chunk, r1 = std::min(rows, r0 + chunk);
        IasDiaDev cb;
        double ms = 0;
        IAS_TRY(ias_dia_mul_dia_rows_dev(&a_dia, &a_dia, r0, r1, &cb, &ms));
        if (ndc == 0) {                                               // first block: the shape of C is known, size the result
            ndc = cb.num_diagonals;
            const size_t cells = (size_t)rows * ndc, cspan = (size_t)std::max(rows + cols - 1, 1);
            o_off = (cells * 8 + 255) / 256 * 256; o_ind = o_off + ((size_t)std::max(ndc, 1) * 4 + 255) / 256 * 256;
            int rc = host_arena(o_ind + cspan * 4 + 256, &base);
            if (rc != IAS_OK) { ias_free_dia_dev(&cb); return rc; }
            h_vals = (double *)base; h_off = (int *)((char *)base + o_off); h_ind = (int *)((char *)base + o_ind);
            if (ndc) IAS_CUDA(cudaMemcpyAsync(h_off, cb.diagonal_offsets_dev, sizeof(int) * (size_t)ndc, cudaMemcpyDeviceToHost, s));
            IAS_TRY(stage[0].alloc((size_t)chunk * std::max(ndc, 1)));
            IAS_TRY(stage[1].alloc((size_t)chunk * std::max(ndc, 1)));
        } else if (cb.num_diagonals != ndc) {
            ias_free_dia_dev(&cb);
            return fail(IAS_E_CUDA, "pipelined DIA path: block %d has %d diagonals, the first had %d", j, cb.num_diagonals, ndc);
        }
        if (j >= 2) IAS_CUDA(cudaStreamWaitEvent(s, ev_d[j - 2], 0));  // the staging buffer's previous block has left
        dia_rows_major(r1 - r0, ndc, cb.values_dev, stage[j & 1].p, s);
        c.launches++;
        IAS_CUDA(cudaEventRecord(ev_c[j], s));
        ias_free_dia_dev(&cb);                                        // stream-ordered: after the transpose
        IAS_CUDA(cudaStreamWaitEvent(s_down, ev_c[j], 0));
        if (r1 > r0 && ndc)
            IAS_CUDA(cudaMemcpyAsync(h_vals + (size_t)r0 * ndc, stage[j & 1].p, sizeof(double) * (size_t)(r1 - r0) * ndc, cudaMemcpyDeviceToHost, s_down));
        IAS_CUDA(cudaEventRecord(ev_d[j], s_down));
        return IAS_OK;
    };
    int rc = IAS_OK;
    for (int k = 0; k < nchunk && rc == IAS_OK; ++k) {
        const int r0 = k * chunk, r1 = std::min(rows, r0 + chunk);
        if ((rc = cudaStreamWaitEvent(s, ev_up[k], 0) == cudaSuccess ? IAS_OK : IAS_E_CUDA) != IAS_OK) break;
        k_flag_diagonals_rows<<<grid_for(r1 - r0, 256), 256, 0, s>>>(r0, r1, rows, rp.p, ci.p, flags_full.p);
        k_fill_dia_rows<<<grid_for(r1 - r0, 256), 256, 0, s>>>(r0, r1, rows, rp.p, ci.p, v.p, flags_spec.p, slot_of.p, a_val.p, bad.p);
        c.launches += 2;
        if (k >= 1) rc = multiply_block(k - 1);
    }
    if (rc == IAS_OK) rc = multiply_block(nchunk - 1);
    if (rc != IAS_OK) { give_up(); *have_dA = 1; return rc; }

    // ---- verification of the speculation, complete features, selection
    // (flags_full == flags_spec entry by entry, no entry fell outside the speculated diagonals)
    {
        k_flags_differ<<<grid_for(span + 1, 256), 256, 0, s>>>(span + 1, flags_full.p, flags_spec.p, bad.p);   // sets bad[1]
        c.launches++;
    }
    int h_bad[2] = {0, 0};
    IAS_CUDA(cudaMemcpyAsync(h_bad, bad.p, sizeof h_bad, cudaMemcpyDeviceToHost, s));
    IAS_CUDA(cudaStreamSynchronize(s));
    if (h_bad[0] || h_bad[1]) return give_up();
    double *f = out->features;
    IAS_TRY(ias_getinfo1(dA, f));
    memcpy(f + 9, f, 9 * sizeof(double));
    ias_getinfo2(rows, cols, nd, f + 18);
    ias_getinfo2(rows, cols, nd, f + 21);
    int wa = 0;
    IAS_TRY(ias_max_row_nnz(dA, &wa));
    ias_getinfo3(rows, nnz, std::max(wa, 1), f + 24);
    ias_getinfo3(rows, nnz, std::max(wa, 1), f + 25);
    const bool ell_ok = ias_sizeof_ell(rows, wa) < gate * ias_sizeof_csr(rows, nnz);
    if (ias_select_format(f, 1, ell_ok ? 1 : 0) != 2) return give_up();
    // diagonal_ind of C (dia:150-158) and the last copies
    {
        const size_t cspan = (size_t)std::max(rows + cols - 1, 1);
        DBuf<int> c_di, c_off;
        IAS_TRY(c_di.alloc(cspan));
        IAS_TRY(c_off.alloc((size_t)std::max(ndc, 1)));
        IAS_CUDA(cudaMemsetAsync(c_di.p, 0, sizeof(int) * cspan, s));
        if (ndc) {
            IAS_CUDA(cudaMemcpyAsync(c_off.p, h_off, sizeof(int) * (size_t)ndc, cudaMemcpyHostToDevice, s));
            IAS_LAUNCH(k_scatter_diag_ind, grid_for(ndc, 256), 256, 0, ndc, rows, c_off.p, c_di.p);
        }
        IAS_CUDA(cudaMemcpyAsync(h_ind, c_di.p, sizeof(int) * cspan, cudaMemcpyDeviceToHost, s));
        IAS_CUDA(cudaStreamSynchronize(s));
        out->d2h_bytes = (long long)((size_t)rows * ndc * 8 + (size_t)ndc * 4 + cspan * 4);
    }
    IAS_CUDA(cudaStreamSynchronize(s_down));
    IAS_CUDA(cudaStreamSynchronize(s_up));
    out->format = 2; out->row = rows; out->col = cols;
    out->values = h_vals; out->diagonal_offsets = h_off; out->diagonal_ind = h_ind;
    out->num_diagonals = ndc; out->nnz = (long long)rows * ndc;
    out->h2d_bytes = (long long)(4 * ((size_t)rows + 1) + 12 * (size_t)nnz);
    *done = 1;
    return IAS_OK;                                                     // rp / ci / v and the DIA operand are released by their holders
}

}  // namespace ias

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
