// spgemm-gpu -- command-line front end of the B200 engine, drop-in for the reference's
// `./spgemm-gpu A.mtx` (GPU/main.cu:30-557) and `./spgemm-cpu A.mtx B.mtx mode` (CPU/main.cpp:97-1003).
//
//   spgemm-gpu A.mtx                 C = A*A   (README.md:10 "All tests default calculate the square of A")
//   spgemm-gpu A.mtx B.mtx [mode]    C = A*B; mode != 0 dumps the operands like testing_mode (CPU/main.cpp:489-497)
//   options: --all (run every format as the reference does), --json, --gate X (default 20, GPU release),
//            --repeat N (best of N timed runs after one warm-up; the reference times one cold run),
//            --opt NAME=VALUE (kernel-selection knob, ias_set_option), --help,
//            --transpose-b (B := A^T, what GPU/main.cu:261-269 computes), --write-c FILE (CSR result as .mtx),
//            --matnet FILE.h5 (MatNet weights; default: ./NetWeights/Intel_weights.h5 and ./NetWeights/P100_weights.h5
//            when present, the paths CPU/MatNet.py:81 and GPU/MatNet.py load), --no-matnet (rule on the features),
//            --stream [GB] (CSR result produced in HBM-budgeted row batches; taken automatically when C does not fit),
//            --gpus N (Algorithm 2 over N GPUs of the box: one process per GPU, every process loads the operands, the rows
//            of A sorted by decreasing products are dealt to the processes in snake order -- ias_row_share -- and each
//            multiplies its share in one streamed pass -- ias_csr_mul_csr_rowlist_stream; no collective on the data path)
//
// Same stages as the reference main: Matrix-Market load -> density images ./imgs/img{1,2}.txt ->
// 26 features -> MatNet format selection -> multiply -> report block (Appendix A of SURVEY.md: run_time,
// trans_time, memory_size, verified_sum, Gflops, Speedup).  Differences, all deliberate (SURVEY
// appendix D): file values are kept (GPU/main.cu:236-243 overwrites them with rand()%10), B := A
// rather than A^T, the selected algorithm is the one that runs (the reference only prints MatNet's pick and
// then runs everything anyway; --all does that), and MatNet.Pred is the engine's native forward pass
// (ias_matnet_*) instead of embedded CPython + Keras.  Without weight files a rule on the feature vector selects.
// Host code only: everything numeric goes through the C ABI of libiaspgemm.so.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../../include/iaspgemm.h"

static int die(const char *what)
{
    fprintf(stderr, "spgemm-gpu: %s: %s\n", what, ias_last_error());
    return 1;
}

static void print_csr(const char *name, const IasCsrMatrix &M)
{
    printf("%s:\n", name);
    for (int i = 0; i < M.row; ++i) {
        for (int p = M.row_ind[i]; p < M.row_ind[i + 1]; ++p) printf("(%d,%d)=%.2f ", i, M.col_ind[p], M.values[p]);
        printf("\n");
    }
    printf("\n");
}

static int write_image(const char *path, const long long *img)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    for (int i = 0; i < 128 * 128; ++i) fprintf(f, "%lld\n", img[i]);
    fclose(f);
    return 0;
}

static bool file_exists(const char *p)
{
    struct stat st;
    return stat(p, &st) == 0 && S_ISREG(st.st_mode);
}

// consumer of a streamed CSR result: appends the batch to an open .mtx file (entries only; the size line was
// written from nnz_total before the first batch)
struct MtxAppend {
    FILE *f;
    bool header_done;
    int rows, cols;
    std::vector<long long> rp;
    std::vector<int> ci;
    std::vector<double> v;
};

static int append_batch(const IasStreamBatch *b, void *user)
{
    MtxAppend *w = (MtxAppend *)user;
    if (!w->f) return 0;
    if (!w->header_done) {
        fprintf(w->f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %lld\n", w->rows, w->cols, b->nnz_total);
        w->header_done = true;
    }
    int n = b->row_end - b->row_begin;
    w->rp.resize((size_t)n + 1); w->ci.resize((size_t)b->batch_nnz); w->v.resize((size_t)b->batch_nnz);
    if (ias_copy(w->rp.data(), b->row_ptr_dev, sizeof(long long) * ((size_t)n + 1), 1)) return IAS_E_CUDA;
    if (b->batch_nnz) {
        if (ias_copy(w->ci.data(), b->col_ind_dev, sizeof(int) * (size_t)b->batch_nnz, 1)) return IAS_E_CUDA;
        if (ias_copy(w->v.data(), b->values_dev, sizeof(double) * (size_t)b->batch_nnz, 1)) return IAS_E_CUDA;
    }
    for (int i = 0; i < n; ++i)
        for (long long p = w->rp[i] - b->entry_base; p < w->rp[i + 1] - b->entry_base; ++p)
            fprintf(w->f, "%d %d %.17g\n", b->row_begin + i + 1, w->ci[p] + 1, w->v[p]);
    return ferror(w->f) ? IAS_E_IO : 0;
}

// ---------------------------------------------------------------- --gpus N: one process per GPU
// What one process reports about its share of the rows of C (sent to the parent through a pipe).
struct ShareResult {
    int rc;                    // 0, or the engine's error code
    int device, rows, batches;
    long long nnz, products;
    double checksum, ms;
    char error[200];
};

// Multiply the rows of `rank` (of `parts`) of A*B on the device this process is bound to: the best of `runs - 1` passes
// after one warm-up (runs > 1), or one cold pass.
static void multiply_share(const IasCsrMatrixDev *dA, const IasCsrMatrixDev *dB, int parts, int rank, size_t budget_bytes, int runs, ShareResult *res)
{
    auto failed = [&](int rc) { res->rc = rc ? rc : IAS_E_CUDA; snprintf(res->error, sizeof(res->error), "%s", ias_last_error()); };
    int *rows_dev = nullptr, count = 0;
    const int cap = (dA->row + parts - 1) / parts;
    int rc = ias_device_alloc((void **)&rows_dev, sizeof(int) * (size_t)(cap > 0 ? cap : 1));
    if (rc) return failed(rc);
    if ((rc = ias_row_share(dA, dB, parts, rank, rows_dev, &count)) != 0) { ias_device_free(rows_dev); return failed(rc); }
    res->rows = count;
    for (int r = 0; r < runs; ++r) {
        IasSpgemmStats st;
        if ((rc = ias_csr_mul_csr_rowlist_stream(dA, dB, rows_dev, count, budget_bytes, nullptr, nullptr, nullptr, &st)) != 0) {
            ias_device_free(rows_dev);
            return failed(rc);
        }
        if (r == (runs > 1 ? 1 : 0) || (r > 0 && st.ms_total < res->ms)) res->ms = st.ms_total;
        res->nnz = st.nnz; res->products = st.products; res->checksum = st.checksum; res->batches = st.batches;
    }
    ias_device_free(rows_dev);
}

// A helper process (rank >= 1): binds to its GPU (a box with fewer devices than ranks shares device 0, so that the
// path can be exercised anywhere), uploads the operands it inherited from the parent, multiplies its share, reports.
static void helper_process(const IasCsrMatrix *A, const IasCsrMatrix *B, bool same, bool transpose_b, int parts, int rank,
                           size_t budget_bytes, int runs, int fd)
{
    ShareResult res;
    memset(&res, 0, sizeof(res));
    res.device = rank;
    auto failed = [&](int rc) { res.rc = rc ? rc : IAS_E_CUDA; snprintf(res.error, sizeof(res.error), "%s", ias_last_error()); };
    IasCsrMatrixDev dA, dB;
    int rc = ias_init(rank);
    if (rc == IAS_E_ARG) { res.device = 0; rc = ias_init(0); }
    if (rc) failed(rc);
    else if ((rc = ias_upload_csr(A, &dA)) != 0) failed(rc);
    else {
        if (transpose_b) rc = ias_csr_transpose(&dA, &dB);
        else if (same) dB = dA;
        else rc = ias_upload_csr(B, &dB);
        if (rc) failed(rc);
        else multiply_share(&dA, &dB, parts, rank, budget_bytes, runs, &res);
    }
    ssize_t w = write(fd, &res, sizeof(res));
    _exit(w == (ssize_t)sizeof(res) && res.rc == 0 ? 0 : 1);            // no stdio flush: the buffers are the parent's
}

int main(int argc, char **argv)
{
    std::vector<std::string> pos;
    bool all = false, json = false, transpose_b = false, no_matnet = false, force_stream = false;
    double stream_gb = 0.0;
    std::string write_c, matnet_path;
    double gate = 20.0;
    int repeat = 1, gpus = 1;
    std::vector<std::pair<std::string, long long> > opts;       // --opt name=value -> ias_set_option
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--help" || a == "-h") {
            printf("usage: spgemm-gpu A.mtx [B.mtx [testing_mode]] [--all] [--json] [--gate X] [--repeat N] [--transpose-b]\n"
                   "                  [--write-c FILE.mtx] [--matnet WEIGHTS.h5 | --no-matnet] [--stream [GB]] [--gpus N] [--opt NAME=VALUE ...]\n"
                   "  --matnet FILE      MatNet weights (default: ./NetWeights/Intel_weights.h5, ./NetWeights/P100_weights.h5 when present)\n"
                   "  --stream [GB]      CSR result in HBM-budgeted row batches (automatic when C does not fit the device)\n"
                   "  --gpus N           Algorithm 2 (CSR) over N GPUs of the box, one process per GPU: rows of A dealt by decreasing work,\n"
                   "                     every process multiplies its share in one streamed pass; report = slowest process, summed checksums\n"
                   "  --opt NAME=VALUE   kernel-selection knob of the engine (ias_set_option, include/iaspgemm.h), e.g.\n"
                   "                     global_rows_smem=0, gwin_max_sw=0; the result does not depend on it\n");
            return 0;
        }
        if (a == "--opt" && i + 1 < argc) {
            std::string kv = argv[++i];
            size_t eq = kv.find('=');
            if (eq == std::string::npos || eq == 0 || eq + 1 >= kv.size()) { printf("--opt expects NAME=VALUE, got '%s'\n", kv.c_str()); return -6; }
            opts.push_back(std::make_pair(kv.substr(0, eq), atoll(kv.c_str() + eq + 1)));
        }
        else if (a == "--all") all = true;
        else if (a == "--json") json = true;
        else if (a == "--transpose-b") transpose_b = true;
        else if (a == "--write-c" && i + 1 < argc) write_c = argv[++i];
        else if (a == "--matnet" && i + 1 < argc) matnet_path = argv[++i];
        else if (a == "--no-matnet") no_matnet = true;
        else if (a == "--stream") {
            force_stream = true;
            if (i + 1 < argc && (argv[i + 1][0] == '.' || (argv[i + 1][0] >= '0' && argv[i + 1][0] <= '9')) && !strstr(argv[i + 1], ".mtx")) stream_gb = atof(argv[++i]);
        }
        else if (a == "--gate" && i + 1 < argc) gate = atof(argv[++i]);
        else if (a == "--repeat" && i + 1 < argc) repeat = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else pos.push_back(a);
    }
    for (size_t o = 0; o < opts.size(); ++o)
        if (ias_set_option(opts[o].first.c_str(), opts[o].second)) { printf("%s\n", ias_last_error()); return -6; }
    if (pos.empty()) {
        printf("please use command like this : ./spgemm-gpu ./sample.mtx\n");
        return -1;
    }
    if (gpus < 1 || gpus > 64) { printf("--gpus expects 1..64, got %d\n", gpus); return -6; }
    if (gpus > 1 && !write_c.empty()) { printf("--write-c needs the whole result in one process: not with --gpus %d\n", gpus); return -6; }
    const std::string fa = pos[0], fb = pos.size() > 1 ? pos[1] : pos[0];
    int testing_mode = pos.size() > 2 ? atoi(pos[2].c_str()) : 0;
    printf("-------------- %s, %s --------------\n", fa.c_str(), fb.c_str());

    IasCsrMatrix A, B;
    int rc = ias_mtx_load(fa.c_str(), &A);
    if (rc) { printf("F1: could not load %s (code %d)\n", fa.c_str(), rc); return rc; }
    bool same = fa == fb;
    const bool host_same = same;               // B's host arrays alias A's
    if (same) B = A;
    else if ((rc = ias_mtx_load(fb.c_str(), &B)) != 0) { printf("F2: could not load %s (code %d)\n", fb.c_str(), rc); return rc; }
    printf("Weight Matrix (A): %dx%d: nnz = %d\n", A.row, A.col, A.nnz);
    printf("Activation Matrix (B): %dx%d: nnz = %d\n", B.row, B.col, B.nnz);
    if (testing_mode) { print_csr("A_csr", A); print_csr("B_csr", B); }
    if (!transpose_b && A.col > B.row) { printf("shape mismatch: A is %dx%d, B is %dx%d\n", A.row, A.col, B.row, B.col); return -5; }

    // --gpus N: the helper processes start here, before this process touches CUDA (a forked child cannot use its parent's
    // context); each one binds to its own GPU and works on its share while this process goes on as rank 0
    const int runs_csr = repeat > 1 ? repeat + 1 : 1;
    std::vector<int> helper_fd;
    std::vector<pid_t> helper_pid;
    fflush(stdout);
    fflush(stderr);
    for (int r = 1; r < gpus; ++r) {
        int fd[2];
        if (pipe(fd) != 0) { printf("pipe failed\n"); return -8; }
        pid_t pid = fork();
        if (pid < 0) { printf("fork failed\n"); return -8; }
        if (pid == 0) {
            close(fd[0]);
            for (size_t k = 0; k < helper_fd.size(); ++k) close(helper_fd[k]);
            helper_process(&A, &B, same, transpose_b, gpus, r, (size_t)(stream_gb * 1e9), runs_csr, fd[1]);
        }
        close(fd[1]);
        helper_fd.push_back(fd[0]);
        helper_pid.push_back(pid);
    }

    if (ias_init(0)) return die("ias_init");
    IasCsrMatrixDev dA, dB;
    if (ias_upload_csr(&A, &dA)) return die("upload A");
    if (transpose_b) {                         // the GPU release's operand: B := A^T
        if (ias_csr_transpose(&dA, &dB)) return die("transpose");
        same = false;
        printf("Activation Matrix (B := A^T): %dx%d: nnz = %d\n", dB.row, dB.col, dB.nnz);
    } else if (same) dB = dA;
    else if (ias_upload_csr(&B, &dB)) return die("upload B");

    // density representation -> ./imgs/img1.txt, ./imgs/img2.txt (CPU/main.cpp:516-643)
    std::vector<long long> img(128 * 128), img_b(128 * 128);
    mkdir("imgs", 0755);
    if (ias_density_image(&dA, img.data())) return die("density A");
    if (write_image("./imgs/img1.txt", img.data())) fprintf(stderr, "spgemm-gpu: cannot write ./imgs/img1.txt\n");
    if (ias_density_image(&dB, img_b.data())) return die("density B");
    if (write_image("./imgs/img2.txt", img_b.data())) fprintf(stderr, "spgemm-gpu: cannot write ./imgs/img2.txt\n");
    printf("------------------------------------------\n");

    double feat[26];
    if (ias_features26(&dA, &dB, feat)) return die("features");
    long long flops = 0;
    if (ias_getflop(&dA, &dB, &flops)) return die("GetFlop");

    // conversions (timed like CPU/main.cpp:658-676: trans_time covers A and B)
    double run[5] = {0}, trans[5] = {0}, size[5] = {0}, sum[5] = {0};
    IasDiaDev a_dia = {}, b_dia = {};
    IasEllDev a_ell = {}, b_ell = {};
    IasCooDev a_coo = {}, b_coo = {};
    struct timespec t0, t1;
    auto now = [](struct timespec *t) { clock_gettime(CLOCK_MONOTONIC, t); };
    auto ms = [](const struct timespec &a, const struct timespec &b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) / 1e6; };
    now(&t0);
    if (ias_csr_to_dia(&dA, gate, &a_dia)) return die("CSRtoDIA");
    if (same) b_dia = a_dia; else if (ias_csr_to_dia(&dB, gate, &b_dia)) return die("CSRtoDIA");
    now(&t1); trans[2] = ms(t0, t1);
    now(&t0);
    if (ias_csr_to_ell(&dA, gate, &a_ell)) return die("CSRtoELL");
    if (same) b_ell = a_ell; else if (ias_csr_to_ell(&dB, gate, &b_ell)) return die("CSRtoELL");
    now(&t1); trans[3] = ms(t0, t1);
    bool dia_ok = a_dia.choice && b_dia.choice, ell_ok = a_ell.choice && b_ell.choice;

    int c = ias_select_format(feat, dia_ok, ell_ok);      // the rule; MatNet overrides it below when weights are present
    // MatNet.Pred (CPU/MatNet.py:24-96, GPU/MatNet.py) on the two density images and the feature vector.  The reference
    // always loads ./NetWeights/<machine>_weights.h5 from the working directory; so does this front end when the files
    // are there.  A 5-class (CPU) net drives the format dispatch; the 3-class (GPU) net names a library SpGEMM, all of
    // which are the engine's CSR pipeline, and its line is printed as GPU/main.cu:543-544 does.
    std::vector<std::string> nets;
    if (!matnet_path.empty()) nets.push_back(matnet_path);
    else if (!no_matnet) {
        if (file_exists("./NetWeights/Intel_weights.h5")) nets.push_back("./NetWeights/Intel_weights.h5");
        if (file_exists("./NetWeights/P100_weights.h5")) nets.push_back("./NetWeights/P100_weights.h5");
    }
    bool dispatched = false;
    for (size_t ni = 0; ni < nets.size(); ++ni) {
        void *net = nullptr;
        if (ias_matnet_load(nets[ni].c_str(), &net)) return die("MatNet weights");
        int nf = 0, nc = 0, cls = 0;
        double probs[8] = {0};
        ias_matnet_shape(net, &nf, &nc, nullptr);
        if (ias_matnet_predict(net, img.data(), img_b.data(), feat, &cls, probs)) return die("MatNet.Pred");
        ias_matnet_free(net);
        if (nc == 5) {                         // CPU nets: 0 MKL 1 CSR 2 DIA 3 ELL 4 COO (the engine runs its CSR path for the MKL slot)
            if (!dispatched) {
                c = cls == 0 ? 1 : cls;
                if ((c == 2 && !dia_ok) || (c == 3 && !ell_ok)) c = 1;      // the size gate refused the format MatNet asked for
                dispatched = true;
            }
        } else {                               // GPU net: CUSP / cuSPARSE / NSPARSE are all CSR SpGEMMs -> the engine's CSR pipeline
            static const char *names[3] = {"CUSP", "cuSPARSE", "NSPARSE"};
            printf("MatNet predicts Algorithm %s is optimal\n", names[cls < 3 ? cls : 0]);
            if (!dispatched && nets.size() == 1) c = 1;
        }
        printf("MatNet class %d of %d (p = %.4f) [%s]\n", cls, nc, probs[cls], nets[ni].c_str());
    }
    printf("The Chosen One = Algorithm %d\n", c + 1);

    auto want = [&](int k) { return all || k == c || (k == 1 && !write_c.empty()); };
    // repeat == 1: one cold run, as the reference times it; repeat > 1: one warm-up, then the best of `repeat`
    const int runs = repeat > 1 ? repeat + 1 : 1;
    auto keep = [&](int r, double t, double &best) { if (r == (runs > 1 ? 1 : 0) || (r > 0 && t < best)) best = t; };
    // Algorithm 2: CSR.  A result that does not fit the device (IAS_E_NOMEM) is produced in HBM-budgeted row batches
    // instead: checksum, nnz and the optional .mtx output come from the batch consumer.
    bool streamed = false;
    int stream_batches = 0;
    if (gpus > 1) {
        // rank 0's share here, the others' from their processes; the report takes the slowest process (what a caller
        // waits for), the summed checksum and the summed nnz
        ShareResult mine;
        memset(&mine, 0, sizeof(mine));
        multiply_share(&dA, &dB, gpus, 0, (size_t)(stream_gb * 1e9), runs_csr, &mine);
        if (mine.rc) { fprintf(stderr, "spgemm-gpu: share 0: %s\n", mine.error); return 1; }
        long long nnz = mine.nnz, products = mine.products;
        run[1] = mine.ms; sum[1] = mine.checksum;
        printf("share 0 on device 0: %d rows, %lld products, nnz %lld, %.3f ms\n", mine.rows, mine.products, mine.nnz, mine.ms);
        for (size_t k = 0; k < helper_fd.size(); ++k) {
            ShareResult res;
            size_t got = 0;
            while (got < sizeof(res)) {
                ssize_t n = read(helper_fd[k], (char *)&res + got, sizeof(res) - got);
                if (n <= 0) break;
                got += (size_t)n;
            }
            close(helper_fd[k]);
            int status = 0;
            waitpid(helper_pid[k], &status, 0);
            if (got != sizeof(res)) { fprintf(stderr, "spgemm-gpu: the process of share %zu ended without a report\n", k + 1); return 1; }
            if (res.rc) { fprintf(stderr, "spgemm-gpu: share %zu: %s\n", k + 1, res.error); return 1; }
            printf("share %zu on device %d: %d rows, %lld products, nnz %lld, %.3f ms\n", k + 1, res.device, res.rows, res.products, res.nnz, res.ms);
            nnz += res.nnz; products += res.products; sum[1] += res.checksum;
            if (res.ms > run[1]) run[1] = res.ms;
        }
        size[1] = ias_sizeof_csr(dA.row, nnz);
        if (products != flops) { fprintf(stderr, "spgemm-gpu: the shares cover %lld products, GetFlop says %lld\n", products, flops); return 1; }
        printf("DONE CSR (rows dealt to %d processes, one per GPU)\n", gpus);
    } else if (want(1)) {
        for (int r = 0; r < runs; ++r) {
            IasCsr64Dev C;
            IasSpgemmStats st;
            int mrc = force_stream ? IAS_E_NOMEM : ias_csr_mul_csr_dev64(&dA, &dB, &C, &st);
            if (mrc == IAS_E_NOMEM) {
                MtxAppend w;
                w.f = nullptr; w.header_done = false; w.rows = dA.row; w.cols = dB.col;
                if (r == runs - 1 && !write_c.empty() && !(w.f = fopen(write_c.c_str(), "w"))) { printf("cannot open %s\n", write_c.c_str()); return -7; }
                if (ias_csr_mul_csr_stream_cb(&dA, &dB, 0, dA.row, (size_t)(stream_gb * 1e9), nullptr, append_batch, &w, &st)) return die("CSR_MUL_CSR_DEV (streamed)");
                if (w.f) {
                    if (!w.header_done) fprintf(w.f, "%%%%MatrixMarket matrix coordinate real general\n%d %d 0\n", w.rows, w.cols);
                    if (fclose(w.f) != 0) { printf("write to %s failed\n", write_c.c_str()); return -7; }
                }
                streamed = true; stream_batches = st.batches;
                keep(r, st.ms_total, run[1]);
                if (r == runs - 1) { sum[1] = st.checksum; size[1] = ias_sizeof_csr(dA.row, st.nnz); }
                continue;
            }
            if (mrc) return die("CSR_MUL_CSR_DEV");
            keep(r, st.ms_total, run[1]);
            if (r == runs - 1) {
                ias_checksum(C.values_dev, C.nnz, &sum[1]); size[1] = ias_sizeof_csr(C.row, C.nnz);
                if (!write_c.empty() && ias_mtx_write_csr64(write_c.c_str(), &C, 0)) return die("write C");
            }
            ias_free_csr64_dev(&C);
        }
        printf(streamed ? "DONE CSR (streamed in %d row batches)\n" : "DONE CSR\n", stream_batches);
    }
    // Algorithm 3: DIA
    if (want(2) && dia_ok) {
        for (int r = 0; r < runs; ++r) {
            IasDiaDev C;
            double t = 0;
            if (ias_dia_mul_dia_dev(&a_dia, &b_dia, &C, &t)) return die("DIA_MUL_DIA_DEV");
            keep(r, t, run[2]);
            if (r == runs - 1) { ias_checksum(C.values_dev, (long long)C.row * C.num_diagonals, &sum[2]); size[2] = ias_sizeof_dia(C.row, C.col, C.num_diagonals); }
            ias_free_dia_dev(&C);
        }
        printf("DONE DIA\n");
    }
    // Algorithm 4: ELL
    if (want(3) && ell_ok) {
        for (int r = 0; r < runs; ++r) {
            IasEllDev C;
            double t = 0;
            if (ias_ell_mul_ell_dev(&a_ell, &b_ell, &C, &t)) return die("ELL_MUL_ELL_DEV");
            keep(r, t, run[3]);
            if (r == runs - 1) { ias_checksum(C.values_dev, (long long)C.row * C.max_nnz_per_row, &sum[3]); size[3] = ias_sizeof_ell(C.row, C.max_nnz_per_row); }
            ias_free_ell_dev(&C);
        }
        printf("DONE ELL\n");
    }
    // Algorithm 5: COO
    if (want(4)) {
        now(&t0);
        if (ias_csr_to_coo(&dA, &a_coo)) return die("CSRtoCOO");
        if (same) b_coo = a_coo; else if (ias_csr_to_coo(&dB, &b_coo)) return die("CSRtoCOO");
        now(&t1); trans[4] = ms(t0, t1);
        for (int r = 0; r < runs; ++r) {
            IasCooDev C;
            double t = 0;
            if (ias_coo_mul_coo_dev(&a_coo, &b_coo, &C, &t)) return die("COO_MUL_COO_DEV");
            keep(r, t, run[4]);
            if (r == runs - 1) { ias_checksum(C.values_dev, C.nnz, &sum[4]); size[4] = ias_sizeof_coo(C.row, C.nnz); }
            ias_free_coo_dev(&C);
        }
        printf("DONE COO\n");
    }

    // report block, CPU/main.cpp:968-1000.  Slot 1 (MKL) is the CPU library row: not run by the GPU front end.
    double base = run[1] > 0 ? run[1] : 0.0, max_speedup = 0.0;
    int max_index = -1;
    double speedup[5];
    for (int i = 0; i < 5; ++i) {
        speedup[i] = (run[i] == 0.0 || base == 0.0) ? 0.0 : base / run[i];
        if (run[i] > 0 && (max_index < 0 || run[i] < run[max_index])) { max_index = i; }
    }
    if (max_index >= 0) max_speedup = base > 0 ? base / run[max_index] : 0.0;
    for (int i = 0; i < 5; ++i) {
        printf("------------------------------\n");
        printf("Algorithm %d:\n", i + 1);
        printf("run_time: %lf\n", run[i]);
        printf("trans_time: %lf\n", trans[i]);
        printf("memory_size: %lf\n", size[i]);
        printf("verified_sum: %lf\n", sum[i]);
        printf("Gflops: %lf\n", run[i] == 0.0 ? 0.0 : (flops * 2.0) / (run[i] * 1000000));
        printf("Speedup: %lf\n", speedup[i]);
    }
    printf("------------------------------\n");
    printf("MAX SPEED IS %lf for ALGORITHM %d\n", max_speedup, max_index + 1);
    printf("------------------------------\n");
    if (all) {
        if (c == max_index) printf("Congratulate! MatNet Correct Prediction.\n");
        else printf("Unfortunately! MatNet Incorrect Prediction.\n");
        printf("------------------------------\n");
    }
    if (json) {
        printf("{\"file_a\": \"%s\", \"file_b\": \"%s\", \"rows\": %d, \"cols\": %d, \"nnz_a\": %d, \"products\": %lld, \"chosen\": %d, \"gpus\": %d, \"features\": [",
               fa.c_str(), fb.c_str(), A.row, dB.col, A.nnz, flops, c + 1, gpus);
        for (int i = 0; i < 26; ++i) printf("%s%.17g", i ? ", " : "", feat[i]);
        printf("], \"run_ms\": [%g, %g, %g, %g, %g], \"verified_sum\": [%.17g, %.17g, %.17g, %.17g, %.17g], \"memory_size\": [%.17g, %.17g, %.17g, %.17g, %.17g]}\n",
               run[0], run[1], run[2], run[3], run[4], sum[0], sum[1], sum[2], sum[3], sum[4], size[0], size[1], size[2], size[3], size[4]);
    }
    ias_free_dia_dev(&a_dia); if (!same) ias_free_dia_dev(&b_dia);
    ias_free_ell_dev(&a_ell); if (!same) ias_free_ell_dev(&b_ell);
    ias_free_coo_dev(&a_coo); if (!same) ias_free_coo_dev(&b_coo);
    ias_free_csr_dev(&dA); if (!same) ias_free_csr_dev(&dB);
    ias_free_host_csr(&A); if (!host_same) ias_free_host_csr(&B);
    return 0;
}
