"""CPU oracle for the shinnosuke training hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

This module is a float64 NumPy *restatement* of the algorithms the reference
(E1eveNn/shinnosuke, pure NumPy) runs on its data-parallel training hot path.
It exists so that the CUDA path in ``shinnosuke_b200`` can be checked against
the reference's arithmetic on a GPU box where ``/root/reference`` does not
exist.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
package never does: it fails loudly when its CUDA library is missing.

Pinning: every function here is checked in ``tests/test_oracle_golden.py``
against golden vectors produced by importing the *live* reference in the build
container (``tests/golden/make_golden.py`` is the generating script; the
``.npz`` fixtures are committed) and against the known-answer values of
SURVEY.md §9.  Parity is therefore pinned to outputs of the reference itself.

Reference defects and the oracle policy for each:
  * ``padding='SAME'`` forward in the reference mis-indexes the padded buffer
    (layers/Convolution.py:137,144).  The oracle implements the true SAME
    convolution, which equals the reference's ``ZeroPadding2D`` + ``VALID``.
  * ``MaxPooling2D`` in 'reshape' mode needs H % pool == 0
    (layers/Convolution.py:319); on odd maps the oracle uses the layer's own
    declared floor semantics (== its 'im2col' mode).
  * Momentum never zeroes gradients (utils/Optimizers.py:52-55): reproduced.

All tensors are float64, C-contiguous, NCHW, exactly like the reference.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# column helpers  (utils/ConvCol.py)
# --------------------------------------------------------------------------

def im2col(x, out_hw, filter_size, stride):
    """Rows ordered (n, i, j), columns ordered (c, u, v); caller pads.
    Follows utils/ConvCol.py:4-21."""
    kh, kw = filter_size
    oh, ow = out_hw
    n, c = x.shape[:2]
    taps = np.empty((n, oh, ow, c, kh, kw), dtype=x.dtype)
    for u in range(kh):
        for v in range(kw):
            win = x[:, :, u:u + stride * oh:stride, v:v + stride * ow:stride]
            taps[:, :, :, :, u, v] = win.transpose(0, 2, 3, 1)
    return taps.reshape(n * oh * ow, c * kh * kw)


def col2im(x_shape, pad, filter_size, stride, dcol):
    """Adjoint scatter-add of :func:`im2col` with padding/crop.
    Follows utils/ConvCol.py:24-42."""
    n, c, h, w = x_shape
    ph, pw = pad
    kh, kw = filter_size
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (w + 2 * pw - kw) // stride + 1
    taps = dcol.reshape(n, oh, ow, c, kh, kw)
    buf = np.zeros((n, c, h + 2 * ph + stride - 1, w + 2 * pw + stride - 1))
    for u in range(kh):
        for v in range(kw):
            buf[:, :, u:u + stride * oh:stride, v:v + stride * ow:stride] += \
                taps[:, :, :, :, u, v].transpose(0, 3, 1, 2)
    return buf[:, :, ph:h + ph, pw:w + pw]


# --------------------------------------------------------------------------
# activations  (utils/Activator.py)
# --------------------------------------------------------------------------

def relu_forward(z):
    """utils/Activator.py:12-15 — output max(0, z); the *pre-activation* is what
    the reference caches for backward."""
    return np.maximum(0, z)


def relu_backward(grad, z):
    """utils/Activator.py:24-28 — strict '<': the gradient passes where z == 0."""
    g = grad.copy()
    g[z < 0] = 0
    return g


def softmax_forward(z):
    """utils/Activator.py:108-110 — the shift is the max over the WHOLE tensor,
    not per row (mathematically the same softmax)."""
    e = np.exp(z - np.max(z))
    return e / np.sum(e, axis=-1, keepdims=True)


def apply_activation(z, act):
    act = (act or 'linear').lower()
    if act == 'relu':
        return relu_forward(z)
    if act == 'softmax':
        return softmax_forward(z)
    if act == 'linear':
        return z
    raise ValueError('oracle: activation %r is off the hot path' % act)


def activation_backward(grad, z, act):
    """Softmax backward is the identity (utils/Activator.py:121-124): the loss
    already supplies (y_hat - y)/N."""
    act = (act or 'linear').lower()
    if act == 'relu':
        return relu_backward(grad, z)
    return grad


# --------------------------------------------------------------------------
# Conv2D  (layers/Convolution.py:131-205)
# --------------------------------------------------------------------------

def conv_output_shape(in_shape, filters, filter_size, stride, padding):
    """layers/Convolution.py:28-50.  Returns (out_shape, pad)."""
    n, c, h, w = in_shape
    kh, kw = filter_size
    padding = padding.upper()
    if padding == 'SAME':
        oh, ow = h, w
        pad = ((stride * (h - 1) - h + kh) // 2, (stride * (w - 1) - w + kw) // 2)
    elif padding == 'VALID':
        oh, ow = (h - kh) // stride + 1, (w - kw) // stride + 1
        pad = (0, 0)
    else:
        raise TypeError('Unknown padding type!plz inputs SAME or VALID')
    return (n, filters, oh, ow), pad


def conv2d_forward(x, w, b, stride=1, pad=(0, 0), act='linear'):
    """Y[n,f,i,j] = act(b[f] + sum_{c,u,v} W[f,c,u,v] * Xp[n,c,i*s+u,j*s+v]).

    layers/Convolution.py:131-161 for VALID; for pad != 0 this is the TRUE
    zero-padded convolution (== reference ZeroPadding2D + VALID), see header.
    Returns (y, cache) where cache keeps what the reference keeps: the column
    matrix and the pre-activation."""
    n, c, h, wd = x.shape
    f, _, kh, kw = w.shape
    ph, pw = pad
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)), 'constant')
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (wd + 2 * pw - kw) // stride + 1
    cols = im2col(xp, (oh, ow), (kh, kw), stride)              # (N*oh*ow, C*kh*kw)
    z = cols.dot(w.reshape(f, -1).T) + np.reshape(b, (1, f))   # (N*oh*ow, F)
    z = np.ascontiguousarray(z.reshape(n, oh, ow, f).transpose(0, 3, 1, 2))
    cache = dict(x_shape=x.shape, cols=cols, z=z, w=w, stride=stride, pad=pad, act=act)
    return apply_activation(z, act), cache


def conv2d_backward(dy, cache, need_dx=True):
    """layers/Convolution.py:182-205: db = sum dY over (n,i,j); dW = dYr . cols^T;
    dX = col2im(dYr^T . Wr).  Returns (dx or None, dw, db) — the caller
    accumulates (+=) like the reference does."""
    w = cache['w']
    f, c, kh, kw = w.shape
    g = activation_backward(dy, cache['z'], cache['act'])
    db = g.sum(axis=(0, 2, 3))
    gr = g.transpose(0, 2, 3, 1).reshape(-1, f)                # rows (n,i,j)
    dw = gr.T.dot(cache['cols']).reshape(f, c, kh, kw)
    dx = None
    if need_dx:
        dcol = gr.dot(w.reshape(f, -1))
        dx = col2im(cache['x_shape'], cache['pad'], (kh, kw), cache['stride'], dcol)
    return dx, dw, db


def conv2d_naive(x, w, b, stride=1, pad=(0, 0)):
    """Independent loop-nest convolution used to cross-check conv2d_forward."""
    n, c, h, wd = x.shape
    f, _, kh, kw = w.shape
    ph, pw = pad
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)), 'constant')
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (wd + 2 * pw - kw) // stride + 1
    y = np.zeros((n, f, oh, ow))
    for i in range(oh):
        for j in range(ow):
            patch = xp[:, :, i * stride:i * stride + kh, j * stride:j * stride + kw]
            y[:, :, i, j] = np.tensordot(patch, w, axes=([1, 2, 3], [1, 2, 3]))
    return y + np.reshape(b, (1, f, 1, 1))


# --------------------------------------------------------------------------
# MaxPooling2D  (layers/Convolution.py:254-362)
# --------------------------------------------------------------------------

POOL_TIE_SPLIT = 'reshape'      # ties share the gradient equally
POOL_FIRST_ARGMAX = 'im2col'    # first maximum (row-major in the window) wins


def maxpool_forward(x, pool=2, stride=None, mode=None):
    """Returns (y, cache).

    mode 'reshape' (layers/Convolution.py:315-324): window max, backward splits
    among ties.  mode 'im2col' (:295-312): ``argmax`` is int64, flat, ordered
    (n, i, j, c), value u*pool+v, first max wins.  Output spatial size is
    floor((H - pool)/stride) + 1 (:283-284); trailing rows/cols are ignored."""
    stride = pool if stride is None else stride
    if mode is None:
        mode = POOL_TIE_SPLIT if pool == stride else POOL_FIRST_ARGMAX
    n, c, h, w = x.shape
    oh, ow = (h - pool) // stride + 1, (w - pool) // stride + 1
    # windows[n,c,i,j,u,v]
    win = np.empty((n, c, oh, ow, pool, pool), dtype=x.dtype)
    for u in range(pool):
        for v in range(pool):
            win[..., u, v] = x[:, :, u:u + stride * oh:stride, v:v + stride * ow:stride]
    flat = win.reshape(n, c, oh, ow, pool * pool)
    y = flat.max(axis=-1)
    cache = dict(x_shape=x.shape, pool=pool, stride=stride, mode=mode)
    if mode == POOL_FIRST_ARGMAX:
        am = flat.argmax(axis=-1)                               # (n,c,i,j)
        cache['argmax'] = np.ascontiguousarray(am.transpose(0, 2, 3, 1)).reshape(-1).astype(np.int64)
    else:
        cache['mask'] = (flat == y[..., None])                  # ties all True
    return y, cache


def maxpool_backward(dy, cache):
    """layers/Convolution.py:333-346 (first-argmax scatter + col2im) and
    :348-362 (tie mask, gradient divided by the number of ties)."""
    n, c, h, w = cache['x_shape']
    pool, stride = cache['pool'], cache['stride']
    oh, ow = dy.shape[2:]
    dx = np.zeros((n, c, h, w))
    if cache['mode'] == POOL_FIRST_ARGMAX:
        am = cache['argmax'].reshape(n, oh, ow, c).transpose(0, 3, 1, 2)
        onehot = (am[..., None] == np.arange(pool * pool))
        contrib = onehot * dy[..., None]
    else:
        mask = cache['mask']
        contrib = mask * (dy / mask.sum(axis=-1))[..., None]
    contrib = contrib.reshape(n, c, oh, ow, pool, pool)
    for u in range(pool):
        for v in range(pool):
            dx[:, :, u:u + stride * oh:stride, v:v + stride * ow:stride] += contrib[..., u, v]
    return dx


# --------------------------------------------------------------------------
# Flatten / Dense  (layers/FC.py)
# --------------------------------------------------------------------------

def flatten_forward(x):
    """layers/FC.py:39-48 — NCHW order (c, i, j) defines the Dense row order."""
    return x.reshape(x.shape[0], -1)


def dense_forward(x, w, b, act='linear'):
    """layers/FC.py:124-134: act(X.W + b)."""
    z = x.dot(w) + b
    return apply_activation(z, act), dict(x=x, w=w, z=z, act=act)


def dense_backward(dy, cache):
    """layers/FC.py:140-153: dW = X^T.dZ, db = sum_rows dZ (keepdims), dX = dZ.W^T
    (always computed).  Returns (dx, dw, db)."""
    g = activation_backward(dy, cache['z'], cache['act'])
    dw = cache['x'].T.dot(g)
    db = g.sum(axis=0, keepdims=True)
    dx = g.dot(cache['w'].T)
    return dx, dw, db


# --------------------------------------------------------------------------
# BatchNormalization  (layers/Normalization.py:93-230), channel axis 1
# --------------------------------------------------------------------------

def batchnorm_forward(x, gamma, beta, moving_mean, moving_var, eps=1e-6, momentum=0.99,
                      is_training=True):
    """Biased variance, eps inside the sqrt, moving stats updated in training
    (layers/Normalization.py:116-132).  Returns (y, cache, new_mean, new_var)."""
    xs = np.moveaxis(x, 1, -1) if x.ndim > 2 else x
    flat = xs.reshape(-1, xs.shape[-1])
    if is_training:
        mean = flat.mean(axis=0)
        var = flat.var(axis=0)
        xmu = flat - mean
        sqrtvar = np.sqrt(var + eps)
        xhat = xmu / sqrtvar
        out = gamma * xhat + beta
        cache = dict(xmu=xmu, sqrtvar=sqrtvar, xhat=xhat, gamma=gamma, x_ndim=x.ndim,
                     moved_shape=xs.shape)
        moving_mean = momentum * moving_mean + (1 - momentum) * mean
        moving_var = momentum * moving_var + (1 - momentum) * var
    else:
        scale = gamma / np.sqrt(moving_var + eps)
        out = flat * scale + (beta - moving_mean * scale)
        cache = None
    out = out.reshape(xs.shape)
    if x.ndim > 2:
        out = np.moveaxis(out, -1, 1)
    return np.ascontiguousarray(out), cache, moving_mean, moving_var


def batchnorm_backward(dy, cache):
    """layers/Normalization.py:186-230.  Returns (dx, dgamma, dbeta)."""
    gs = np.moveaxis(dy, 1, -1) if dy.ndim > 2 else dy
    g = gs.reshape(-1, gs.shape[-1])
    xmu, sqrtvar, xhat, gamma = cache['xmu'], cache['sqrtvar'], cache['xhat'], cache['gamma']
    dbeta = g.sum(axis=0)
    dgamma = (g * xhat).sum(axis=0)
    m = xhat.shape[0]
    dxhat = g * gamma
    dvar = np.sum(np.power(-1. / sqrtvar, 3) * xmu * dxhat * 0.5, axis=0)
    dmean = np.sum(-dxhat / sqrtvar, axis=0) - 2 * dvar * np.mean(xmu, axis=0)
    dx = dxhat / sqrtvar + dvar * 2 * xmu / m + dmean / m
    dx = dx.reshape(gs.shape)
    if dy.ndim > 2:
        dx = np.moveaxis(dx, -1, 1)
    return np.ascontiguousarray(dx), dgamma, dbeta


# --------------------------------------------------------------------------
# Embedding (layers/Embeddings.py:54-76) and LSTM (layers/Recurrent.py:216-437), BASELINE config 5
# --------------------------------------------------------------------------

def embedding_forward(w, idx):
    """layers/Embeddings.py:60: output = W[input]."""
    return w[np.asarray(idx, dtype=np.int64)]


def embedding_backward(w_shape, idx, dy):
    """layers/Embeddings.py:75-76: np.add.at(W.grads, input, grads)."""
    dw = np.zeros(w_shape)
    np.add.at(dw, np.asarray(idx, dtype=np.int64), dy)
    return dw


def _sigmoid(z):
    return 1 / (1 + np.exp(-z))        # utils/Activator.py:58


def lstm_forward(x, w, b, h0=None, c0=None):
    """LSTMCell._forward, layers/Recurrent.py:265-318.  x (N,T,D); w (H+D, 4H) rows [recurrent;
    input], gate columns (f, u, c~, o); b (1,4H).  Returns h for every step (N,T,H) and the cache."""
    n, t_steps, d = x.shape
    h_units = w.shape[1] // 4
    h = np.zeros((n, t_steps + 1, h_units)); c = np.zeros((n, t_steps + 1, h_units))
    if h0 is not None:
        h[:, 0] = h0
    if c0 is not None:
        c[:, 0] = c0
    f_a = np.zeros((n, t_steps, h_units)); u_a = np.zeros_like(f_a); o_a = np.zeros_like(f_a); ct_a = np.zeros_like(f_a)
    z = np.zeros((n, t_steps, d + h_units))
    for t in range(1, t_steps + 1):
        zt = np.concatenate((h[:, t - 1], x[:, t - 1]), axis=1)          # :294
        ot = zt.dot(w) + b                                                # :295
        f = _sigmoid(ot[:, :h_units]); u = _sigmoid(ot[:, h_units:2 * h_units])
        ct = np.tanh(ot[:, 2 * h_units:3 * h_units]); o = _sigmoid(ot[:, 3 * h_units:])
        c[:, t] = f * c[:, t - 1] + u * ct                                # :301
        h[:, t] = o * np.tanh(c[:, t])                                    # :302
        f_a[:, t - 1], u_a[:, t - 1], o_a[:, t - 1], ct_a[:, t - 1] = f, u, o, ct
        z[:, t - 1] = zt
    return h[:, 1:], dict(z=z, h=h, c=c, f=f_a, u=u_a, o=o_a, ct=ct_a, w=w, units=h_units)


def lstm_backward(pre_grad, cache, return_sequences, reference_gate_order=False):
    """LSTMCell._backward, layers/Recurrent.py:321-437.  Returns (dx, dw, db).

    The reference concatenates the gate gradients as (o, c~, u, f) (:358, :406) although the forward
    columns of W are (f, u, c~, o) (:296-299): its dW, db and dX are therefore not the gradient of
    its own forward pass.  reference_gate_order=True reproduces the reference bit for bit (pinned by
    tests/golden/lstm.npz); the default is the correct order, pinned by numeric differentiation in
    tests/test_oracle_golden.py — that is what the device implements."""
    z, h, c, w, hu = cache['z'], cache['h'], cache['c'], cache['w'], cache['units']
    n, t_steps, _ = z.shape
    dw = np.zeros_like(w); db = np.zeros((1, 4 * hu))
    dx = np.zeros((n, t_steps, z.shape[2] - hu))
    da_next = np.zeros((n, hu)); dc_next = np.zeros((n, hu))
    for t in reversed(range(t_steps)):
        if return_sequences:
            da = pre_grad[:, t] + da_next                                  # :327
        else:
            da = pre_grad if t == t_steps - 1 else da_next                 # :372, :416
        tc = np.tanh(c[:, t + 1])
        f, u, o, ct = cache['f'][:, t], cache['u'][:, t], cache['o'][:, t], cache['ct'][:, t]
        do = da * tc * o * (1 - o)                                         # :328-329
        dc = dc_next + da * o * (1 - np.square(tc))                        # :336-337
        dct = dc * u * (1 - np.square(ct))                                 # :339-340
        du = dc * ct * u * (1 - u)                                         # :346-347
        df = dc * c[:, t] * f * (1 - f)                                    # :352-353
        dgrad = np.concatenate((do, dct, du, df) if reference_gate_order else (df, du, dct, do), axis=1)   # :358
        dw += z[:, t].T.dot(dgrad)                                         # :360
        db += dgrad.sum(axis=0, keepdims=True)                             # :362
        dz = dgrad.dot(w.T)                                                # :364
        da_next = dz[:, :hu]
        dc_next = dc * f                                                   # :367
        dx[:, t] = dz[:, hu:]
    return dx, dw, db


# --------------------------------------------------------------------------
# loss  (utils/Objectives.py:72-91) — one-hot labels despite the "Sparse" name
# --------------------------------------------------------------------------

def scce_loss(y_hat, y):
    avg = np.prod(np.asarray(y_hat.shape[:-1]))
    return float(-np.sum(y * np.log(y_hat)) / avg)


def scce_acc(y_hat, y):
    return float(np.mean(np.argmax(y_hat, axis=-1) == np.argmax(y, axis=-1)))


def scce_backward(y_hat, y):
    avg = np.prod(np.asarray(y_hat.shape[:-1]))
    return (y_hat - y) / avg


# --------------------------------------------------------------------------
# optimizers  (utils/Optimizers.py:24-157)
# --------------------------------------------------------------------------

def sgd_update(params, grads, lr=0.01):
    """w <- w - lr*g, g <- 0 (utils/Optimizers.py:29-32).  In-place on the lists."""
    for i in range(len(params)):
        params[i] = params[i] - lr * grads[i]
        grads[i] = np.zeros_like(params[i])


def momentum_update(params, grads, velocity, lr=0.01, rho=0.9):
    """v <- rho*v + (1-rho)*g ; w <- w - lr*v ; g is NOT zeroed
    (utils/Optimizers.py:46-55)."""
    for i in range(len(params)):
        velocity[i] = rho * velocity[i] + (1 - rho) * grads[i]
        params[i] = params[i] - lr * velocity[i]


def rmsprop_update(params, grads, ms, lr=0.001, rho=0.9, eps=1e-7):
    """s <- rho*s + (1-rho)*g^2 ; w <- w - lr*g/sqrt(s+eps) ; g NOT zeroed (utils/Optimizers.py:72-80)."""
    for i in range(len(params)):
        ms[i] = rho * ms[i] + (1 - rho) * np.square(grads[i])
        params[i] = params[i] - lr * grads[i] / np.sqrt(ms[i] + eps)


def adagrad_update(params, grads, ms, lr=0.01, eps=1e-7):
    """s <- s + g^2 ; w <- w - lr*g/sqrt(s+eps) ; g NOT zeroed (utils/Optimizers.py:90-97)."""
    for i in range(len(params)):
        ms[i] = ms[i] + np.power(grads[i], 2)
        params[i] = params[i] - lr * grads[i] / np.sqrt(ms[i] + eps)


def adadelta_update(params, grads, ms, dx, rho=0.95, eps=1e-7):
    """s <- rho*s + (1-rho)*g^2 ; u = sqrt((x+eps)/(s+eps))*g ; w <- w - u ; x <- rho*x + (1-rho)*u^2.
    `lr` is not used; g NOT zeroed (utils/Optimizers.py:110-123)."""
    for i in range(len(params)):
        ms[i] = rho * ms[i] + (1 - rho) * np.power(grads[i], 2)
        u = np.sqrt((dx[i] + eps) / (ms[i] + eps)) * grads[i]
        params[i] = params[i] - u
        dx[i] = rho * dx[i] + (1 - rho) * np.power(u, 2)


def adam_update(params, grads, vs, ms, iterations, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-7):
    """utils/Optimizers.py:138-157.  `iterations` is the counter value AFTER Adam's own `+= 1`
    (the base class adds another one afterwards, so successive updates see 1, 3, 5, ...):
    v <- b1*v + (1-b1)*g ; m <- b2*m + (1-b2)*g^2 ;
    w <- w - lr * (v/(1-b1^it)) / (sqrt(m/(1-b2^it)) + eps) ; g NOT zeroed."""
    for i in range(len(params)):
        vs[i] = beta1 * vs[i] + (1 - beta1) * grads[i]
        ms[i] = beta2 * ms[i] + (1 - beta2) * np.square(grads[i])
        v_c = vs[i] / (1 - pow(beta1, iterations))
        m_c = ms[i] / (1 - pow(beta2, iterations))
        params[i] = params[i] - lr * (v_c / (np.sqrt(m_c) + eps))


# --------------------------------------------------------------------------
# mini-batching  (utils/MiniBatch.py:6-34)
# --------------------------------------------------------------------------

def get_batches(x, y, batch_size, seed, shuffle):
    """Reseeds the GLOBAL NumPy RNG with the epoch index, permutes, slices; the
    last partial batch is kept."""
    np.random.seed(seed)
    m = x.shape[0]
    if shuffle:
        perm = np.random.permutation(m)
        x, y = x[perm], y[perm]
    out = []
    for s in range(0, m, batch_size):
        out.append((x[s:s + batch_size], y[s:s + batch_size]))
    return out


def to_categorical(labels):
    """utils/Preprocess.py:5-20."""
    labels = np.asarray(labels).reshape(-1)
    return np.eye(int(labels.max()) + 1)[labels]


# --------------------------------------------------------------------------
# a tiny sequential driver (models.py:30-137) used for trajectory parity and as
# the bench CPU baseline.  ``spec`` is a list of tuples:
#   ('conv', F, k, stride, padding, act) | ('pool', p) | ('pool', p, mode) |
#   ('flatten',) | ('dense', n_out, act) | ('bn',) | ('act', name)
#   ('embedding', vocab, dim) | ('lstm', units, return_sequences)      (BASELINE config 5)
#   ('tap', key) | ('add', key)  — residual connection (layers/Base.py:352-411, Add): 'tap'
#   remembers the current map, 'add' sums it back in; backward sends the gradient both ways.
# --------------------------------------------------------------------------

class OracleNet:
    def __init__(self, spec, input_shape, optimizer='sgd', lr=0.01, rho=0.9, std=0.1):
        """Parameter creation consumes the global NumPy stream in the reference's
        order: Conv2D W ~ Normal(std 0.1) then b ~ randn(1,F)
        (layers/Convolution.py:52-53); Dense W ~ Normal(std 0.1), b zeros
        (layers/FC.py:107-108); BN gamma ones, beta zeros."""
        self.spec = list(spec)
        self.optimizer, self.lr, self.rho = optimizer, lr, rho
        self.params, self.grads, self.state = [], [], []
        self.slots = []            # per layer: indices into params
        shape = tuple(input_shape)
        for op in self.spec:
            kind = op[0]
            if kind == 'conv':
                _, f, k, stride, padding, act = op
                w = np.random.normal(0.0, std, size=(f, shape[1], k, k))
                b = np.random.randn(1, f)
                self.slots.append(self._add(w, b))
                shape, pad = conv_output_shape(shape, f, (k, k), stride, padding)
                self.state.append(dict(pad=pad))
            elif kind == 'pool':
                p = op[1]
                shape = (shape[0], shape[1], (shape[2] - p) // p + 1, (shape[3] - p) // p + 1)
                self.slots.append(()); self.state.append({})
            elif kind == 'flatten':
                shape = (shape[0], int(np.prod(shape[1:])))
                self.slots.append(()); self.state.append({})
            elif kind == 'dense':
                _, n_out, act = op
                w = np.random.normal(0.0, std, size=(shape[-1], n_out))
                b = np.zeros((1, n_out))
                self.slots.append(self._add(w, b))
                shape = shape[:-1] + (n_out,)
                self.state.append({})
            elif kind == 'bn':
                c = shape[1]
                self.slots.append(self._add(np.ones(c), np.zeros(c)))
                self.state.append(dict(mm=np.zeros(c), mv=np.ones(c)))
            elif kind in ('act', 'tap', 'add'):
                self.slots.append(()); self.state.append({})
            elif kind == 'embedding':
                _, vocab, dim = op
                self.slots.append(self._add(np.random.uniform(-0.05, 0.05, size=(vocab, dim))))      # Initializers.py:29-35
                shape = tuple(shape) + (dim,)
                self.state.append({})
            elif kind == 'lstm':
                # draw order of layers/Recurrent.py:229-240: per gate (f, u, c~, o) the input block
                # (glorot uniform), then the recurrent block (orthogonal: SVD of a standard normal matrix)
                _, units, rs = op
                d = shape[-1]
                cols = []
                for _g in range(4):
                    fan = int(np.sqrt(d * units))          # utils/Initializers.py:12-22: shape tuples always take the sqrt(prod) branch
                    lim = np.sqrt(6. / (fan + fan))
                    w_l = np.random.uniform(-lim, lim, size=(d, units))
                    u_, _s, v_ = np.linalg.svd(np.random.normal(loc=0.0, scale=1.0, size=(units, units)), full_matrices=False)
                    cols.append(np.concatenate((u_, w_l), axis=0))
                b = np.zeros((1, 4 * units)); b[:, :units] = 1.0                                     # unit_forget_bias (:251-254)
                self.slots.append(self._add(np.concatenate(cols, axis=1), b))
                shape = (shape[0], shape[1], units) if rs else (shape[0], units)
                self.state.append({})
            else:
                raise ValueError(kind)
        self.velocity = None
        self.train_loss, self.train_acc = [], []

    def _add(self, *arrs):
        idx = []
        for a in arrs:
            idx.append(len(self.params))
            self.params.append(a)
            self.grads.append(np.zeros_like(a))
        return tuple(idx)

    def forward(self, x, is_training=True):
        caches = []
        taps = {}
        h = x
        for op, slot, st in zip(self.spec, self.slots, self.state):
            kind = op[0]
            if kind == 'conv':
                _, f, k, stride, padding, act = op
                h, c = conv2d_forward(h, self.params[slot[0]], self.params[slot[1]], stride, st['pad'], act)
            elif kind == 'pool':
                mode = op[2] if len(op) > 2 else None
                if mode is None and (h.shape[2] % op[1] or h.shape[3] % op[1]):
                    mode = POOL_FIRST_ARGMAX     # header: floor semantics on odd maps
                h, c = maxpool_forward(h, op[1], None, mode)
            elif kind == 'flatten':
                c = h.shape
                h = flatten_forward(h)
            elif kind == 'dense':
                h, c = dense_forward(h, self.params[slot[0]], self.params[slot[1]], op[2])
            elif kind == 'bn':
                h, c, mm, mv = batchnorm_forward(h, self.params[slot[0]], self.params[slot[1]],
                                                 st['mm'], st['mv'], is_training=is_training)
                if is_training:
                    st['mm'], st['mv'] = mm, mv
            elif kind == 'act':
                c = h
                h = apply_activation(h, op[1])
            elif kind == 'tap':
                taps[op[1]] = h
                c = None
            elif kind == 'add':
                h = taps[op[1]] + h
                c = None
            elif kind == 'embedding':
                c = h
                h = embedding_forward(self.params[slot[0]], h)
            elif kind == 'lstm':
                hs, c = lstm_forward(h, self.params[slot[0]], self.params[slot[1]])
                h = hs if op[2] else hs[:, -1]
            caches.append(c)
        return h, caches

    def backward(self, dy, caches):
        g = dy
        skip = {}
        for li in range(len(self.spec) - 1, -1, -1):
            op, slot, c = self.spec[li], self.slots[li], caches[li]
            kind = op[0]
            if kind == 'conv':
                g, dw, db = conv2d_backward(g, c, need_dx=(li > 0))
                self.grads[slot[0]] = self.grads[slot[0]] + dw
                self.grads[slot[1]] = self.grads[slot[1]] + db.reshape(self.grads[slot[1]].shape)
            elif kind == 'pool':
                g = maxpool_backward(g, c)
            elif kind == 'flatten':
                g = g.reshape(c)
            elif kind == 'dense':
                g, dw, db = dense_backward(g, c)
                self.grads[slot[0]] = self.grads[slot[0]] + dw
                self.grads[slot[1]] = self.grads[slot[1]] + db
            elif kind == 'bn':
                g, dgm, dbt = batchnorm_backward(g, c)
                self.grads[slot[0]] = self.grads[slot[0]] + dgm
                self.grads[slot[1]] = self.grads[slot[1]] + dbt
            elif kind == 'act':
                g = activation_backward(g, c, op[1])
            elif kind == 'add':
                skip[op[1]] = g                     # the same gradient reaches both operands
            elif kind == 'tap':
                g = g + skip[op[1]]
            elif kind == 'embedding':
                self.grads[slot[0]] = self.grads[slot[0]] + embedding_backward(self.params[slot[0]].shape, c, g)
            elif kind == 'lstm':
                g, dw, db = lstm_backward(g, c, op[2])
                self.grads[slot[0]] = self.grads[slot[0]] + dw
                self.grads[slot[1]] = self.grads[slot[1]] + db

    def step(self, xs, ys):
        """One minibatch of models.py:70-92: forward, loss grad, backward sweep,
        optimizer update; loss/acc are from the pre-update y_hat."""
        y_hat, caches = self.forward(xs, True)
        self.backward(scce_backward(y_hat, ys), caches)
        if self.optimizer == 'sgd':
            sgd_update(self.params, self.grads, self.lr)
        else:
            if self.velocity is None:
                self.velocity = [np.zeros_like(p) for p in self.params]
            momentum_update(self.params, self.grads, self.velocity, self.lr, self.rho)
        loss, acc = scce_loss(y_hat, ys), scce_acc(y_hat, ys)
        self.train_loss.append(loss)
        self.train_acc.append(acc)
        return loss, acc

    def fit(self, x, y, batch_size, epochs=1, shuffle=True):
        for epoch in range(epochs):
            for xs, ys in get_batches(x, y, batch_size, epoch, shuffle):
                self.step(xs, ys)


# Config builders shared by tests and bench (BASELINE.json `configs`).
def spec_cfg1():
    return [('conv', 8, 2, 1, 'VALID', 'relu'), ('pool', 2), ('flatten',), ('dense', 10, 'softmax')], (None, 1, 28, 28)


def spec_cfg2():
    return [('dense', 500, 'relu'), ('dense', 10, 'softmax')], (None, 784)


def spec_residual(width, k=3, padding='SAME', stem=True):
    """One residual block of BASELINE config 4: [Conv-BN-ReLU-Conv-BN] + skip, ReLU, 2x2 pool."""
    s = [('conv', width, k, 1, padding, 'relu')] if stem else []
    s += [('tap', 'skip'), ('conv', width, k, 1, padding, 'linear'), ('bn',), ('act', 'relu'),
          ('conv', width, k, 1, padding, 'linear'), ('bn',), ('add', 'skip'), ('act', 'relu'), ('pool', 2)]
    return s


def spec_cfg4():
    """stem Conv64 -> residual block @64 -> pool -> Conv128 -> residual block @128 -> pool -> Dense10."""
    s = spec_residual(64) + spec_residual(128) + [('flatten',), ('dense', 10, 'softmax')]
    return s, (None, 3, 32, 32)


def spec_cfg5(vocab=1000, dim=64, units=128, timesteps=64, classes=2):
    """ints [0, vocab) of shape (N, 64) -> Embedding 64 -> LSTM 128 -> Dense 2 (softmax)."""
    return [('embedding', vocab, dim), ('lstm', units, False), ('dense', classes, 'softmax')], (None, timesteps)


def spec_cfg3():
    s = []
    for f in (64, 128, 256, 256):
        s += [('conv', f, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2)]
    s += [('flatten',), ('dense', 10, 'softmax')]
    return s, (None, 3, 32, 32)
