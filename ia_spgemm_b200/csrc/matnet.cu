// matnet.cu -- the MatNet format selector, natively (host only): weight loading + forward pass.
//
// Replaces the embedded-CPython call of the reference (CPU/main.cpp:682-704, GPU/main.cu:446-460) into
// CPU/MatNet.py:24-96, which rebuilds a Keras model on every call, loads ./NetWeights/*.h5 and predicts
// the fastest algorithm from the two 128x128 density images and the feature vector.  Neither Keras,
// TensorFlow nor h5py exist in this image, so this file carries
//   (1) a minimal HDF5 reader for exactly what Keras 2.1 writes: superblock v0, version-1 object headers,
//       version-1 group B-trees + local heaps + symbol-table nodes, contiguous little-endian float32
//       datasets (the weight files are ~219 KB, 20 tensors);
//   (2) the network of MatNet.py:45-79 in fp32: two towers Conv3x3(16,valid,tanh) - MaxPool2 -
//       Conv5x5(16,stride 2,same,tanh) - MaxPool2 - Conv5x5(16,stride 2,same,tanh) - MaxPool2 - Flatten(256) -
//       Dense32 tanh; features -> Dense(F) tanh; concat(32+32+F) -> Dense(classes) softmax; argmax.
//       Images are rescaled img*255/max (MatNet.py:29-37); TensorFlow "same" padding (extra pad at the end).
// Layer names follow Keras' creation order in MatNet.py: conv2d_1..3 tower 1, conv2d_4..6 tower 2,
// dense_1 features, dense_2 / dense_3 the towers' Dense32, dense_4 the classifier.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct Tensor {
    std::vector<long long> dims;
    std::vector<float> data;
};

struct MatNet {
    std::map<std::string, Tensor> t;       // "conv2d_1/kernel", "dense_4/bias", ...
    int n_features = 0, n_classes = 0;
};

// ---------------------------------------------------------------- minimal HDF5
struct H5 {
    std::vector<unsigned char> b;
    unsigned long long base = 0;
    bool ok(size_t off, size_t n) const { return off <= b.size() && n <= b.size() - off; }     // no wrap-around
    unsigned char at(size_t off) const { return off < b.size() ? b[off] : 0; }
    template <class T>
    T rd(size_t off) const
    {
        T v{};
        if (ok(off, sizeof(T))) memcpy(&v, b.data() + off, sizeof(T));
        return v;
    }
    bool sig(size_t off, const char *s) const { return ok(off, 4) && memcmp(b.data() + off, s, 4) == 0; }

    struct Msg { int type; size_t body; int size; };
    bool messages(size_t addr, std::vector<Msg> &out) const
    {
        if (!ok(addr, 16) || b[addr] != 1) return false;           // version-1 object header
        int nmsg = rd<unsigned short>(addr + 2);
        unsigned hsize = rd<unsigned>(addr + 8);
        std::vector<std::pair<size_t, size_t>> blocks{{addr + 16, hsize}};
        for (size_t bi = 0; bi < blocks.size() && (int)out.size() < nmsg; ++bi) {
            size_t pos = blocks[bi].first, end = pos + blocks[bi].second;
            while (pos + 8 <= end && (int)out.size() < nmsg) {
                int type = rd<unsigned short>(pos), size = rd<unsigned short>(pos + 2);
                size_t body = pos + 8;
                if (!ok(body, size)) return false;
                if (type == 0x10) blocks.push_back({(size_t)rd<unsigned long long>(body), (size_t)rd<unsigned long long>(body + 8)});
                out.push_back({type, body, size});
                pos = body + size;
            }
        }
        return true;
    }
    bool symbol_table(size_t addr, size_t &btree, size_t &heap) const
    {
        std::vector<Msg> m;
        if (!messages(addr, m)) return false;
        for (const Msg &x : m)
            if (x.type == 0x11) { btree = rd<unsigned long long>(x.body); heap = rd<unsigned long long>(x.body + 8); return true; }
        return false;
    }
    void walk(size_t node, size_t names, std::vector<std::pair<std::string, size_t>> &out, int depth = 0) const
    {
        if (depth > 16) return;
        if (sig(node, "TREE")) {
            int used = rd<unsigned short>(node + 6);
            size_t pos = node + 24;
            for (int i = 0; i < used; ++i, pos += 16) walk((size_t)rd<unsigned long long>(pos + 8), names, out, depth + 1);
        } else if (sig(node, "SNOD")) {
            int n = rd<unsigned short>(node + 6);
            size_t pos = node + 8;
            for (int i = 0; i < n; ++i, pos += 40) {
                size_t s = names + (size_t)rd<unsigned long long>(pos);
                size_t e = s;
                while (e < b.size() && b[e]) ++e;
                out.push_back({std::string((const char *)b.data() + s, e - s), (size_t)rd<unsigned long long>(pos + 8)});
            }
        }
    }
    bool links(size_t addr, std::vector<std::pair<std::string, size_t>> &out) const
    {
        size_t btree, heap;
        if (!symbol_table(addr, btree, heap) || !sig(heap, "HEAP")) return false;
        walk(btree, (size_t)rd<unsigned long long>(heap + 24), out);
        return true;
    }
    bool dataset(size_t addr, Tensor &t) const
    {
        std::vector<Msg> m;
        if (!messages(addr, m)) return false;
        bool f32 = false, have_dims = false;
        size_t data = 0, size = 0;
        for (const Msg &x : m) {
            if (x.type == 0x01) {
                if (x.size < 2) return false;
                int ver = at(x.body), rank = at(x.body + 1);
                if (rank > 8) return false;
                size_t off = x.body + (ver == 1 ? 8 : 4);
                t.dims.clear();
                for (int r = 0; r < rank; ++r) t.dims.push_back((long long)rd<unsigned long long>(off + 8 * r));
                have_dims = true;
            } else if (x.type == 0x03) {
                f32 = x.size >= 8 && (at(x.body) & 0x0F) == 1 && rd<unsigned>(x.body + 4) == 4 && (at(x.body + 1) & 1) == 0;
            } else if (x.type == 0x08) {
                if (x.size < 18 || at(x.body) != 3 || at(x.body + 1) != 1) return false;     // layout v3, contiguous
                data = (size_t)rd<unsigned long long>(x.body + 2);
                size = (size_t)rd<unsigned long long>(x.body + 10);
            }
        }
        if (!f32 || !have_dims || !size) return false;
        size_t n = 1;
        for (long long d : t.dims) {
            if (d < 0 || (d > 0 && n > b.size() / (size_t)d)) return false;      // more elements than the file has bytes
            n *= (size_t)d;
        }
        if (size != 4 * n || data > b.size() || !ok(base + data, size)) return false;
        t.data.resize(n);
        memcpy(t.data.data(), b.data() + base + data, size);
        return true;
    }
    void collect(size_t addr, const std::string &prefix, std::map<std::string, Tensor> &out, int depth = 0) const
    {
        std::vector<std::pair<std::string, size_t>> l;
        if (depth > 8 || !links(addr, l)) return;
        for (auto &kv : l) {
            size_t bt, hp;
            if (symbol_table(kv.second, bt, hp)) collect(kv.second, prefix.empty() ? kv.first : prefix + "/" + kv.first, out, depth + 1);
            else {
                Tensor t;
                if (dataset(kv.second, t)) out[prefix + "/" + kv.first] = t;
            }
        }
    }
};

// ---------------------------------------------------------------- layers (fp32, HWC images, HWIO kernels)
void conv2d_tanh(const std::vector<float> &in, int H, int W, int Cin, const Tensor &k, const Tensor &bias, int stride, bool same,
                 std::vector<float> &out, int &Ho, int &Wo)
{
    int kh = (int)k.dims[0], kw = (int)k.dims[1], Cout = (int)k.dims[3];
    int pt = 0, pl = 0;
    if (same) {
        Ho = (H + stride - 1) / stride; Wo = (W + stride - 1) / stride;
        int ph = std::max((Ho - 1) * stride + kh - H, 0), pw = std::max((Wo - 1) * stride + kw - W, 0);
        pt = ph / 2; pl = pw / 2;                     // TensorFlow puts the odd padding row/column at the end
    } else {
        Ho = (H - kh) / stride + 1; Wo = (W - kw) / stride + 1;
    }
    out.assign((size_t)Ho * Wo * Cout, 0.f);
    for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Wo; ++x) {
            float *o = &out[((size_t)y * Wo + x) * Cout];
            for (int c = 0; c < Cout; ++c) o[c] = bias.data[c];
            for (int dy = 0; dy < kh; ++dy) {
                int iy = y * stride + dy - pt;
                if (iy < 0 || iy >= H) continue;
                for (int dx = 0; dx < kw; ++dx) {
                    int ix = x * stride + dx - pl;
                    if (ix < 0 || ix >= W) continue;
                    const float *ip = &in[((size_t)iy * W + ix) * Cin];
                    const float *kp = &k.data[(((size_t)dy * kw + dx) * Cin) * Cout];
                    for (int ci = 0; ci < Cin; ++ci) {
                        float v = ip[ci];
                        const float *kc = kp + (size_t)ci * Cout;
                        for (int c = 0; c < Cout; ++c) o[c] += v * kc[c];
                    }
                }
            }
            for (int c = 0; c < Cout; ++c) o[c] = tanhf(o[c]);
        }
}

void maxpool2(const std::vector<float> &in, int H, int W, int C, std::vector<float> &out, int &Ho, int &Wo)
{
    Ho = H / 2; Wo = W / 2;
    out.assign((size_t)Ho * Wo * C, 0.f);
    for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Wo; ++x)
            for (int c = 0; c < C; ++c) {
                float m = in[((size_t)(2 * y) * W + 2 * x) * C + c];
                m = fmaxf(m, in[((size_t)(2 * y) * W + 2 * x + 1) * C + c]);
                m = fmaxf(m, in[((size_t)(2 * y + 1) * W + 2 * x) * C + c]);
                m = fmaxf(m, in[((size_t)(2 * y + 1) * W + 2 * x + 1) * C + c]);
                out[((size_t)y * Wo + x) * C + c] = m;
            }
}

void dense(const std::vector<float> &in, const Tensor &k, const Tensor &bias, bool tanh_act, std::vector<float> &out)
{
    int n_in = (int)k.dims[0], n_out = (int)k.dims[1];
    out.assign(n_out, 0.f);
    for (int o = 0; o < n_out; ++o) {
        float s = bias.data[o];
        for (int i = 0; i < n_in; ++i) s += in[i] * k.data[(size_t)i * n_out + o];
        out[o] = tanh_act ? tanhf(s) : s;
    }
}

const Tensor *need(const MatNet &n, const char *name, int rank)
{
    auto it = n.t.find(name);
    if (it == n.t.end() || (int)it->second.dims.size() != rank) return nullptr;
    return &it->second;
}

bool tower(const MatNet &net, const long long *img, int first_conv, const char *dense_name, std::vector<float> &out)
{
    long long mx = 0;
    for (int i = 0; i < 128 * 128; ++i) mx = img[i] > mx ? img[i] : mx;
    std::vector<float> a(128 * 128), b;
    for (int i = 0; i < 128 * 128; ++i) a[i] = mx > 0 ? (float)((double)img[i] * 255.0 / (double)mx) : 0.f;
    int H = 128, W = 128, C = 1;
    for (int l = 0; l < 3; ++l) {
        char kn[64], bn[64];
        snprintf(kn, sizeof kn, "conv2d_%d/kernel", first_conv + l);
        snprintf(bn, sizeof bn, "conv2d_%d/bias", first_conv + l);
        const Tensor *k = need(net, kn, 4), *bi = need(net, bn, 1);
        if (!k || !bi || k->dims[2] != C) return false;
        int Ho, Wo;
        conv2d_tanh(a, H, W, C, *k, *bi, l == 0 ? 1 : 2, l != 0, b, Ho, Wo);
        C = (int)k->dims[3];
        maxpool2(b, Ho, Wo, C, a, H, W);
    }
    std::string dk = std::string(dense_name) + "/kernel", db = std::string(dense_name) + "/bias";
    const Tensor *k = need(net, dk.c_str(), 2), *bi = need(net, db.c_str(), 1);
    if (!k || !bi || k->dims[0] != (long long)a.size()) return false;
    dense(a, *k, *bi, true, out);
    return true;
}

// "conv2d_1/conv2d_1/kernel:0" -> "conv2d_1/kernel"
std::string short_name(const std::string &full)
{
    size_t s = full.rfind('/');
    std::string leaf = s == std::string::npos ? full : full.substr(s + 1);
    std::string rest = s == std::string::npos ? "" : full.substr(0, s);
    size_t s2 = rest.rfind('/');
    std::string layer = s2 == std::string::npos ? rest : rest.substr(s2 + 1);
    size_t c = leaf.find(':');
    if (c != std::string::npos) leaf = leaf.substr(0, c);
    return layer + "/" + leaf;
}

}  // namespace

using namespace ias;

extern "C" {

int ias_matnet_create(void **net)
{
    if (!net) return fail(IAS_E_ARG, "NULL");
    *net = new MatNet();
    return IAS_OK;
}

void ias_matnet_free(void *net) { delete static_cast<MatNet *>(net); }

int ias_matnet_set_tensor(void *net, const char *name, const float *data, const long long *dims, int rank)
{
    if (!net || !name || !data || !dims || rank < 1 || rank > 4) return fail(IAS_E_ARG, "bad tensor");
    Tensor t;
    size_t n = 1;
    for (int r = 0; r < rank; ++r) { t.dims.push_back(dims[r]); n *= (size_t)dims[r]; }
    t.data.assign(data, data + n);
    MatNet *m = static_cast<MatNet *>(net);
    m->t[name] = t;
    if (!strcmp(name, "dense_1/kernel")) m->n_features = (int)dims[0];
    if (!strcmp(name, "dense_4/kernel")) m->n_classes = (int)dims[1];
    return IAS_OK;
}

int ias_matnet_get_tensor(void *net, const char *name, float *out, long long *dims, int *rank)
{
    if (!net || !name) return fail(IAS_E_ARG, "NULL");
    MatNet *m = static_cast<MatNet *>(net);
    auto it = m->t.find(name);
    if (it == m->t.end()) return fail(IAS_E_ARG, "no tensor named %s", name);
    if (rank) *rank = (int)it->second.dims.size();
    if (dims) for (size_t r = 0; r < it->second.dims.size(); ++r) dims[r] = it->second.dims[r];
    if (out) memcpy(out, it->second.data.data(), sizeof(float) * it->second.data.size());
    return IAS_OK;
}

int ias_matnet_load(const char *path, void **net)
{
    if (!path || !net) return fail(IAS_E_ARG, "NULL");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(IAS_E_IO, "cannot open %s", path);
    H5 h;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    h.b.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = fread(h.b.data(), 1, h.b.size(), f);
    fclose(f);
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (got != h.b.size() || h.b.size() < 96 || memcmp(h.b.data(), magic, 8) != 0 || h.b[8] != 0 || h.b[13] != 8 || h.b[14] != 8)
        return fail(IAS_E_IO, "%s is not an HDF5 file with a version-0 superblock and 8-byte offsets", path);
    h.base = h.rd<unsigned long long>(24);
    size_t root = (size_t)h.rd<unsigned long long>(24 + 32 + 8);
    std::map<std::string, Tensor> all;
    h.collect(root, "", all);
    MatNet *m = new MatNet();
    for (auto &kv : all) {
        std::string n = short_name(kv.first);
        m->t[n] = kv.second;
    }
    const Tensor *d1 = need(*m, "dense_1/kernel", 2), *d4 = need(*m, "dense_4/kernel", 2);
    if (!d1 || !d4 || m->t.size() < 20) { delete m; return fail(IAS_E_IO, "%s does not hold the 20 MatNet tensors", path); }
    m->n_features = (int)d1->dims[0];
    m->n_classes = (int)d4->dims[1];
    *net = m;
    return IAS_OK;
}

int ias_matnet_shape(void *net, int *n_features, int *n_classes, long long *n_params)
{
    if (!net) return fail(IAS_E_ARG, "NULL");
    MatNet *m = static_cast<MatNet *>(net);
    if (n_features) *n_features = m->n_features;
    if (n_classes) *n_classes = m->n_classes;
    if (n_params) { long long p = 0; for (auto &kv : m->t) p += (long long)kv.second.data.size(); *n_params = p; }
    return IAS_OK;
}

// MatNet.Pred (CPU/MatNet.py:24-96): class index 0..classes-1 and (optionally) the softmax probabilities
int ias_matnet_predict(void *net, const long long *img1, const long long *img2, const double *features, int *cls, double *probs)
{
    if (!net || !img1 || !img2 || !features || !cls) return fail(IAS_E_ARG, "NULL");
    MatNet *m = static_cast<MatNet *>(net);
    std::vector<float> t1, t2, fd, cat, logits;
    if (!tower(*m, img1, 1, "dense_2", t1) || !tower(*m, img2, 4, "dense_3", t2)) return fail(IAS_E_ARG, "MatNet tower weights missing or mis-shaped");
    const Tensor *k1 = need(*m, "dense_1/kernel", 2), *b1 = need(*m, "dense_1/bias", 1);
    const Tensor *k4 = need(*m, "dense_4/kernel", 2), *b4 = need(*m, "dense_4/bias", 1);
    if (!k1 || !b1 || !k4 || !b4) return fail(IAS_E_ARG, "MatNet dense weights missing");
    std::vector<float> fin(m->n_features);
    for (int i = 0; i < m->n_features; ++i) fin[i] = (float)features[i];
    dense(fin, *k1, *b1, true, fd);
    cat = t1; cat.insert(cat.end(), t2.begin(), t2.end()); cat.insert(cat.end(), fd.begin(), fd.end());   // Concatenate([image1, image2, features])
    if ((long long)cat.size() != k4->dims[0]) return fail(IAS_E_ARG, "MatNet classifier expects %lld inputs, got %zu", k4->dims[0], cat.size());
    dense(cat, *k4, *b4, false, logits);
    float mx = logits[0];
    for (float v : logits) mx = fmaxf(mx, v);
    double sum = 0.0;
    std::vector<double> e(logits.size());
    for (size_t i = 0; i < logits.size(); ++i) { e[i] = exp((double)(logits[i] - mx)); sum += e[i]; }
    int best = 0;
    for (size_t i = 0; i < logits.size(); ++i) {
        if (probs) probs[i] = e[i] / sum;
        if (logits[i] > logits[best]) best = (int)i;
    }
    *cls = best;
    return IAS_OK;
}

}  // extern "C"
