"""NumPy restatement of MatNet's forward pass (CPU/MatNet.py:24-96, Keras 2.1 / TensorFlow semantics) in
float64 -- TEST HELPER used to pin the engine's native fp32 implementation (csrc/matnet.cu)."""
import numpy as np


def conv2d(x, k, b, stride, same):
    H, W, Cin = x.shape
    kh, kw, _, Cout = k.shape
    if same:
        Ho, Wo = -(-H // stride), -(-W // stride)
        ph, pw = max((Ho - 1) * stride + kh - H, 0), max((Wo - 1) * stride + kw - W, 0)
        x = np.pad(x, ((ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))      # TF: the odd pad goes to the end
    else:
        Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    out = np.zeros((Ho, Wo, Cout))
    for dy in range(kh):
        for dx in range(kw):
            patch = x[dy:dy + (Ho - 1) * stride + 1:stride, dx:dx + (Wo - 1) * stride + 1:stride, :]
            out += patch @ k[dy, dx]
    return np.tanh(out + b)


def maxpool2(x):
    H, W, C = x.shape
    return x[:H // 2 * 2, :W // 2 * 2].reshape(H // 2, 2, W // 2, 2, C).max(axis=(1, 3))


def tower(img, w, first, dense):
    x = np.asarray(img, dtype=np.float64).reshape(128, 128)
    mx = x.max()
    x = (x * 255.0 / mx if mx > 0 else x)[:, :, None]
    for l in range(3):
        x = conv2d(x, w["conv2d_%d/kernel" % (first + l)], w["conv2d_%d/bias" % (first + l)], 1 if l == 0 else 2, l != 0)
        x = maxpool2(x)
    return np.tanh(x.reshape(-1) @ w[dense + "/kernel"] + w[dense + "/bias"])


def predict(w, img1, img2, features):
    w = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
    nf = w["dense_1/kernel"].shape[0]
    f = np.tanh(np.asarray(features, dtype=np.float64)[:nf] @ w["dense_1/kernel"] + w["dense_1/bias"])
    cat = np.concatenate([tower(img1, w, 1, "dense_2"), tower(img2, w, 4, "dense_3"), f])
    logits = cat @ w["dense_4/kernel"] + w["dense_4/bias"]
    e = np.exp(logits - logits.max())
    return int(np.argmax(logits)), e / e.sum()


def random_weights(n_features, n_classes, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    w = {}
    for i, cin in ((1, 1), (2, 16), (3, 16), (4, 1), (5, 16), (6, 16)):
        k = 3 if cin == 1 else 5
        w["conv2d_%d/kernel" % i] = rng.normal(0, scale * (0.3 if cin == 1 else 0.08), size=(k, k, cin, 16))
        w["conv2d_%d/bias" % i] = rng.normal(0, 0.05, size=16)
    for name, (a, b) in {"dense_1": (n_features, n_features), "dense_2": (256, 32), "dense_3": (256, 32),
                         "dense_4": (64 + n_features, n_classes)}.items():
        w[name + "/kernel"] = rng.normal(0, scale * 0.2, size=(a, b))
        w[name + "/bias"] = rng.normal(0, 0.05, size=b)
    return {k: v.astype(np.float32) for k, v in w.items()}
