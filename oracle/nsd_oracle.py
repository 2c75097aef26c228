"""CPU oracle for the GRUDecoder + CTC hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the arithmetic the reference performs
on the hot path.  It is the *checker* for the CUDA
kernels: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``neural_speech_decoder_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests and no golden vectors ("parity
unpinned" by the reference's own suite).  This oracle is therefore pinned
against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference/src``) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
every function below against those fixtures.

The arithmetic of the path lives in third-party torch (pinned ``torch==1.13.1``
in the reference's ``setup.cfg:43``; 2.11.0 in this image): ``F.conv1d``,
``torch.einsum``, ``nn.Softsign``, ``nn.Unfold``, ``nn.GRU``, ``nn.Linear``,
``log_softmax``, ``nn.CTCLoss``, ``argmax``/``unique_consecutive``.  Each
function cites the reference call site it restates.

All functions are dtype-generic: pass float64 arrays for a high-precision
check, float32 to mimic the reference's working precision.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

NEG_INF = -np.inf


# --------------------------------------------------------------------------
# a1  Gaussian smoothing            (augmentations.py:41-69 taps, :83-91 conv)
# --------------------------------------------------------------------------
def gaussian_taps(sigma: float, ntaps: int = 20) -> np.ndarray:
    """Normalised Gaussian FIR taps in float32, built the way
    ``GaussianSmoothing.__init__`` does (augmentations.py:50-63): float32 grid,
    mean=(n-1)/2, 1/(s*sqrt(2pi))*exp(-((g-mean)/s)^2/2), divided by its sum."""
    if sigma == 0:
        raise ZeroDivisionError("gaussianSmoothWidth must be > 0 (augmentations.py:57)")
    grid = np.arange(ntaps, dtype=np.float32)
    mean = np.float32((ntaps - 1) / 2)
    k = np.float32(1.0 / (sigma * math.sqrt(2 * math.pi))) * np.exp(
        -(((grid - mean) / np.float32(sigma)) ** 2) / np.float32(2)
    ).astype(np.float32)
    k = (k / k.sum(dtype=np.float32)).astype(np.float32)
    return k


def smooth(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """Depthwise FIR over time with ``padding="same"`` (augmentations.py:91,
    called model.py:84-86).  For an even tap count torch pads
    left=(n-1)//2, right=n-1-left, i.e. left 9 / right 10 for 20 taps:
    y[b,t,c] = sum_k w[k] * x[b, t-left+k, c], zero outside [0,T).
    x: [B,T,N]."""
    B, T, N = x.shape
    n = taps.shape[0]
    left = (n - 1) // 2
    xp = np.zeros((B, T + n - 1, N), dtype=x.dtype)
    xp[:, left:left + T, :] = x
    y = np.zeros_like(x)
    for k in range(n):
        y += taps[k].astype(x.dtype) * xp[:, k:k + T, :]
    return y


# --------------------------------------------------------------------------
# a2/a3  day affine + softsign                         (model.py:89-93)
# --------------------------------------------------------------------------
def day_affine(y: np.ndarray, day_w: np.ndarray, day_b: np.ndarray, day_idx: np.ndarray) -> np.ndarray:
    """pre[b,t,k] = sum_d y[b,t,d] * dayWeights[dayIdx[b],d,k] + dayBias[dayIdx[b],0,k]
    (einsum "btd,bdk->btk", model.py:89-92)."""
    day_idx = np.asarray(day_idx)
    if day_idx.size and (day_idx.min() < 0 or day_idx.max() >= day_w.shape[0]):
        raise IndexError("dayIdx out of range")
    W = day_w[day_idx]              # [B,N,N]
    b = day_b[day_idx]              # [B,1,N]
    return np.matmul(y, W) + b


def softsign(x: np.ndarray) -> np.ndarray:
    """x / (1 + |x|)  (model.py:36, :93)."""
    return x / (1 + np.abs(x))


# --------------------------------------------------------------------------
# a4  unfold                                             (model.py:37-39, 96-101)
# --------------------------------------------------------------------------
def n_frames(T: int, kernel_len: int, stride_len: int) -> int:
    if T < kernel_len:
        raise RuntimeError("sequence shorter than kernelLen")
    return (T - kernel_len) // stride_len + 1


def unfold(z: np.ndarray, kernel_len: int, stride_len: int) -> np.ndarray:
    """patches[b,j,c*K+k] = z[b, j*S+k, c]  (channel-major, tap-minor).
    z: [B,T,N] -> [B,T',N*K]."""
    B, T, N = z.shape
    Tp = n_frames(T, kernel_len, stride_len)
    out = np.empty((B, Tp, N, kernel_len), dtype=z.dtype)
    for j in range(Tp):
        win = z[:, j * stride_len:j * stride_len + kernel_len, :]      # [B,K,N]
        out[:, j] = np.transpose(win, (0, 2, 1))
    return out.reshape(B, Tp, N * kernel_len)


def unfold_bwd(dp: np.ndarray, T: int, kernel_len: int, stride_len: int) -> np.ndarray:
    """col2im: dz[b,t,c] = sum_{(j,k): j*S+k=t} dp[b,j,c*K+k]."""
    B, Tp, F = dp.shape
    N = F // kernel_len
    d4 = dp.reshape(B, Tp, N, kernel_len)
    dz = np.zeros((B, T, N), dtype=dp.dtype)
    for j in range(Tp):
        dz[:, j * stride_len:j * stride_len + kernel_len, :] += np.transpose(d4[:, j], (0, 2, 1))
    return dz


def frontend_fwd(x, taps, day_w, day_b, day_idx, kernel_len, stride_len):
    """model.py:84-101 in one call.  Returns (patches, saved) with the
    intermediates the backward needs."""
    ys = smooth(x, taps)
    pre = day_affine(ys, day_w, day_b, day_idx)
    z = softsign(pre)
    patches = unfold(z, kernel_len, stride_len)
    return patches, {"ys": ys, "pre": pre, "z": z}


def frontend_bwd(dpatches, saved, day_w, day_idx, T, kernel_len, stride_len):
    """Gradients of dayWeights / dayBias (the input X has no grad).
    dW[d] = sum_{b:dayIdx_b=d} ys_b^T dpre_b ; db[d] = sum_b sum_t dpre_b."""
    dz = unfold_bwd(dpatches, T, kernel_len, stride_len)
    pre = saved["pre"]
    dpre = dz / (1 + np.abs(pre)) ** 2
    ys = saved["ys"]
    d_day_w = np.zeros_like(day_w, dtype=dpatches.dtype)
    d_day_b = np.zeros((day_w.shape[0], 1, day_w.shape[1]), dtype=dpatches.dtype)
    for b, d in enumerate(np.asarray(day_idx)):
        d_day_w[d] += ys[b].T @ dpre[b]
        d_day_b[d, 0] += dpre[b].sum(axis=0)
    return d_day_w, d_day_b


# --------------------------------------------------------------------------
# a5/a6  GRU                                            (model.py:50-63, 119)
# --------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_dir_fwd(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """One layer, one direction, h0 = 0 (model.py:104-117).  x: [B,T',in].
    ``w_ih is None``: x already IS the input projection gi [B,T',3H] (recurrence-only checks at sizes where an
    identity W_ih would cost a 3H x 3H matmul per row).
    Gate row order [r,z,n]:
      gi = x W_ih^T + b_ih ; gh = h_{t-1} W_hh^T + b_hh
      r = s(gi_r+gh_r) ; z = s(gi_z+gh_z) ; n = tanh(gi_n + r*gh_n)
      h_t = (1-z)*n + z*h_{t-1}
    The reverse direction walks t = T'-1 .. 0 over the full padded length.
    Returns (hseq [B,T',H], saved)."""
    B, Tp, _ = x.shape
    H = w_hh.shape[1]
    gi = x if w_ih is None else x @ w_ih.T + b_ih      # time-batched input projection (a5)
    hseq = np.zeros((B, Tp, H), dtype=x.dtype)
    r_s = np.zeros_like(hseq); z_s = np.zeros_like(hseq)
    n_s = np.zeros_like(hseq); hn_s = np.zeros_like(hseq); hp_s = np.zeros_like(hseq)
    h = np.zeros((B, H), dtype=x.dtype)
    order = range(Tp - 1, -1, -1) if reverse else range(Tp)
    for t in order:
        gh = h @ w_hh.T + b_hh
        r = _sigmoid(gi[:, t, :H] + gh[:, :H])
        z = _sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
        hp_s[:, t] = h
        h = (1 - z) * n + z * h
        hseq[:, t] = h
        r_s[:, t] = r; z_s[:, t] = z; n_s[:, t] = n; hn_s[:, t] = gh[:, 2 * H:]
    return hseq, {"x": x, "r": r_s, "z": z_s, "n": n_s, "hn": hn_s, "hprev": hp_s}


def gru_dir_bwd(dhseq, saved, w_ih, w_hh, reverse=False):
    """BPTT for one layer-direction.  dhseq: [B,T',H] gradient w.r.t. every
    emitted h_t.  Returns dx, dw_ih, dw_hh, db_ih, db_hh."""
    x = saved["x"]
    B, Tp, _ = x.shape
    H = w_hh.shape[1]
    dgi = np.zeros((B, Tp, 3 * H), dtype=x.dtype)
    dgh = np.zeros((B, Tp, 3 * H), dtype=x.dtype)
    dh = np.zeros((B, H), dtype=x.dtype)
    order = range(Tp) if reverse else range(Tp - 1, -1, -1)   # opposite of forward
    for t in order:
        r, z, n, hn, hp = (saved[k][:, t] for k in ("r", "z", "n", "hn", "hprev"))
        dht = dhseq[:, t] + dh
        dn = dht * (1 - z)
        dz = dht * (hp - n)
        dnt = dn * (1 - n * n)
        dzt = dz * z * (1 - z)
        drt = dnt * hn * r * (1 - r)
        dgi[:, t] = np.concatenate([drt, dzt, dnt], axis=1)
        dgh[:, t] = np.concatenate([drt, dzt, dnt * r], axis=1)
        dh = dht * z + dgh[:, t] @ w_hh
    dgi2 = dgi.reshape(B * Tp, 3 * H)
    dgh2 = dgh.reshape(B * Tp, 3 * H)
    if w_ih is None:                            # x was gi itself: dx = dgi, no input weights
        dw_hh = dgh2.T @ saved["hprev"].reshape(B * Tp, H)
        return dgi, None, dw_hh, dgi2.sum(0), dgh2.sum(0)
    dx = (dgi2 @ w_ih).reshape(B, Tp, -1)
    dw_ih = dgi2.T @ x.reshape(B * Tp, -1)
    dw_hh = dgh2.T @ saved["hprev"].reshape(B * Tp, H)
    return dx, dw_ih, dw_hh, dgi2.sum(0), dgh2.sum(0)


def gru_fwd(x, weights: Sequence[Dict[str, np.ndarray]], bidirectional: bool):
    """Stacked GRU, eval mode (no dropout).  ``weights[l]`` holds
    ``w_ih,w_hh,b_ih,b_hh`` and, if bidirectional, the ``*_reverse`` set
    (torch naming: weight_ih_l{l}[_reverse]).  Layer l>0 input = cat(fwd,bwd)."""
    saved = []
    inp = x
    for lw in weights:
        hf, sf = gru_dir_fwd(inp, lw["w_ih"], lw["w_hh"], lw["b_ih"], lw["b_hh"], False)
        if bidirectional:
            hb, sb = gru_dir_fwd(inp, lw["w_ih_reverse"], lw["w_hh_reverse"],
                                 lw["b_ih_reverse"], lw["b_hh_reverse"], True)
            inp = np.concatenate([hf, hb], axis=2)
            saved.append((sf, sb))
        else:
            inp = hf
            saved.append((sf,))
    return inp, saved


def gru_bwd(dout, saved, weights, bidirectional: bool):
    grads: List[Dict[str, np.ndarray]] = [dict() for _ in weights]
    d = dout
    for l in range(len(weights) - 1, -1, -1):
        lw = weights[l]
        H = lw["w_hh"].shape[1]
        dx, dwi, dwh, dbi, dbh = gru_dir_bwd(d[:, :, :H], saved[l][0], lw["w_ih"], lw["w_hh"], False)
        grads[l].update(w_ih=dwi, w_hh=dwh, b_ih=dbi, b_hh=dbh)
        if bidirectional:
            dx2, dwi, dwh, dbi, dbh = gru_dir_bwd(d[:, :, H:], saved[l][1], lw["w_ih_reverse"],
                                                  lw["w_hh_reverse"], True)
            grads[l].update(w_ih_reverse=dwi, w_hh_reverse=dwh, b_ih_reverse=dbi, b_hh_reverse=dbh)
            dx = dx + dx2
        d = dx
    return d, grads


# --------------------------------------------------------------------------
# a8/a9  output projection, log-softmax            (model.py:122; trainer:210)
# --------------------------------------------------------------------------
def linear(x, w, b):
    return x @ w.T + b


def log_softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    s = x - m
    return s - np.log(np.exp(s).sum(axis=axis, keepdims=True))


# --------------------------------------------------------------------------
# a10  out_lens                                             (trainer:209, 300)
# --------------------------------------------------------------------------
def out_lens(x_len: np.ndarray, kernel_len: int, stride_len: int) -> np.ndarray:
    """((X_len - K) / S).to(int32): true division then truncation toward 0."""
    return np.trunc((np.asarray(x_len).astype(np.float32) - kernel_len) / np.float32(stride_len)).astype(np.int32)


# --------------------------------------------------------------------------
# a11  CTC loss (blank=0, reduction="mean", zero_infinity=True)
#      (trainer:139-141, 213-218; arithmetic = torch ctc_loss)
# --------------------------------------------------------------------------
def _lse2(a, b):
    if a == NEG_INF and b == NEG_INF:
        return NEG_INF
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def _lse3(a, b, c):
    m = max(a, b, c)
    if m == NEG_INF:
        return NEG_INF
    return m + math.log(math.exp(a - m) + math.exp(b - m) + math.exp(c - m))


def ctc_alpha_beta(lp: np.ndarray, target: np.ndarray, in_len: int, tgt_len: int, blank: int = 0):
    """lp: [T,C] log-probs of one utterance.  Returns (nll, log_alpha, log_beta)
    over the blank-extended lattice S'=2*tgt_len+1.  A skip s-2 -> s is allowed
    iff ext[s] != ext[s-2] (which also excludes blanks)."""
    S = 2 * tgt_len + 1
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = target[:tgt_len]
    la = np.full((max(in_len, 1), S), NEG_INF, dtype=np.float64)
    lb = np.full((max(in_len, 1), S), NEG_INF, dtype=np.float64)
    if in_len <= 0:
        # torch: with zero input frames only the empty target is feasible
        return (0.0 if tgt_len == 0 else np.inf), la, lb
    la[0, 0] = lp[0, blank]
    if S > 1:
        la[0, 1] = lp[0, ext[1]]
    for t in range(1, in_len):
        for s in range(S):
            a1 = la[t - 1, s]
            a2 = la[t - 1, s - 1] if s > 0 else NEG_INF
            a3 = la[t - 1, s - 2] if (s > 1 and ext[s - 2] != ext[s]) else NEG_INF
            v = _lse3(a1, a2, a3)
            la[t, s] = v + lp[t, ext[s]] if v != NEG_INF else NEG_INF
    l1 = la[in_len - 1, S - 1]
    l2 = la[in_len - 1, S - 2] if S > 1 else NEG_INF
    nll = -_lse2(l1, l2)
    lb[in_len - 1, S - 1] = lp[in_len - 1, blank]
    if S > 1:
        lb[in_len - 1, S - 2] = lp[in_len - 1, ext[S - 2]]
    for t in range(in_len - 2, -1, -1):
        for s in range(S):
            b1 = lb[t + 1, s]
            b2 = lb[t + 1, s + 1] if s < S - 1 else NEG_INF
            b3 = lb[t + 1, s + 2] if (s < S - 2 and ext[s + 2] != ext[s]) else NEG_INF
            v = _lse3(b1, b2, b3)
            lb[t, s] = v + lp[t, ext[s]] if v != NEG_INF else NEG_INF
    return nll, la, lb


def ctc_loss(log_probs: np.ndarray, targets: np.ndarray, input_lengths, target_lengths,
             blank: int = 0, reduction: str = "mean", zero_infinity: bool = True):
    """log_probs: [T,B,C].  targets: [B,Smax] padded.  Returns
    (loss, per-utterance nll [B], grad wrt log_probs [T,B,C]) where ``grad`` is
    what torch's ctc backward returns for ``loss.backward()``:
    (exp(lp) - exp(log sum_{s:ext[s]=c}(alpha+beta) + nll - lp)) * gscale for
    t < input_length, 0 beyond; gscale = 1/(clamp(tgt_len,1)*B) for "mean",
    1 for "sum"/"none" (a "none" grad assumes upstream grad of ones)."""
    T, B, C = log_probs.shape
    nll = np.zeros(B, dtype=np.float64)
    grad = np.zeros((T, B, C), dtype=np.float64)
    lp64 = log_probs.astype(np.float64)
    for b in range(B):
        il, tl = int(input_lengths[b]), int(target_lengths[b])
        n, la, lb = ctc_alpha_beta(lp64[:, b], targets[b], il, tl, blank)
        if zero_infinity and (n == np.inf or np.isnan(n)):
            nll[b] = 0.0
            continue
        nll[b] = n
        gs = 1.0 / (max(tl, 1) * B) if reduction == "mean" else 1.0
        S = 2 * tl + 1
        ext = np.full(S, blank, dtype=np.int64); ext[1::2] = targets[b][:tl]
        for t in range(il):
            acc = np.full(C, NEG_INF)
            for s in range(S):
                acc[ext[s]] = _lse2(acc[ext[s]], la[t, s] + lb[t, s])
            with np.errstate(over="ignore", invalid="ignore"):
                occ = np.where(acc == NEG_INF, 0.0, np.exp(acc + n - lp64[t, b]))
            grad[t, b] = (np.exp(lp64[t, b]) - occ) * gs
    if reduction == "mean":
        tl = np.maximum(np.asarray(target_lengths, dtype=np.float64), 1.0)
        loss = float(np.mean(nll / tl))
    elif reduction == "sum":
        loss = float(nll.sum())
    else:
        loss = nll.copy()
    return loss, nll, grad


# --------------------------------------------------------------------------
# a12/a13  greedy decode + edit distance                  (trainer:313-333)
# --------------------------------------------------------------------------
def greedy_decode(log_probs: np.ndarray, lens: np.ndarray, blank: int = 0) -> List[List[int]]:
    """log_probs: [T,B,C].  argmax (ties -> lowest index) over frames
    t < lens[b], collapse consecutive repeats, drop blanks."""
    T, B, C = log_probs.shape
    out = []
    for b in range(B):
        L = max(0, min(int(lens[b]), T))
        ids = np.argmax(log_probs[:L, b, :], axis=-1) if L > 0 else np.zeros(0, dtype=np.int64)
        seq, prev = [], None
        for v in ids.tolist():
            if v != prev:
                if v != blank:
                    seq.append(int(v))
                prev = v
        out.append(seq)
    return out


def edit_distance(a: Sequence[int], b: Sequence[int]) -> int:
    """Levenshtein distance with unit sub/ins/del: what
    ``edit_distance.SequenceMatcher(a,b).distance()`` returns (trainer:322-330)."""
    la, lb = len(a), len(b)
    prev = list(range(lb + 1))
    for i in range(1, la + 1):
        cur = [i] + [0] * lb
        for j in range(1, lb + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[i - 1] != b[j - 1]))
        prev = cur
    return prev[lb]


def phoneme_error_rate(decoded: Sequence[Sequence[int]], targets: np.ndarray, target_lengths) -> Tuple[int, int]:
    """(sum of edit distances, sum of true lengths); cer = dist/len (trainer:332-333)."""
    dist = tot = 0
    for b, dec in enumerate(decoded):
        true = [int(v) for v in targets[b][:int(target_lengths[b])]]
        dist += edit_distance(true, list(dec))
        tot += len(true)
    return dist, tot


# --------------------------------------------------------------------------
# whole model                                              (model.py:83-123)
# --------------------------------------------------------------------------
def split_gru_state(sd: Dict[str, np.ndarray], n_layers: int, bidirectional: bool):
    """Pick the nn.GRU tensors out of a GRUDecoder state dict."""
    ws = []
    for l in range(n_layers):
        d = {}
        for sfx, key in (("", ""), ("_reverse", "_reverse")) if bidirectional else (("", ""),):
            d["w_ih" + key] = sd[f"gru_decoder.weight_ih_l{l}{sfx}"]
            d["w_hh" + key] = sd[f"gru_decoder.weight_hh_l{l}{sfx}"]
            d["b_ih" + key] = sd[f"gru_decoder.bias_ih_l{l}{sfx}"]
            d["b_hh" + key] = sd[f"gru_decoder.bias_hh_l{l}{sfx}"]
        ws.append(d)
    return ws


def decoder_forward(sd: Dict[str, np.ndarray], x, day_idx, *, kernel_len, stride_len, n_layers,
                    bidirectional, dtype=np.float64):
    """GRUDecoder.forward (eval / dropout 0).  ``sd`` = state dict as numpy.
    Returns (logits [B,T',C], saved-for-backward)."""
    c = lambda a: np.asarray(a).astype(dtype)
    taps = c(sd["gaussianSmoother.weight"][0, 0])
    patches, fsaved = frontend_fwd(c(x), taps, c(sd["dayWeights"]), c(sd["dayBias"]), day_idx,
                                   kernel_len, stride_len)
    ws = [{k: c(v) for k, v in d.items()} for d in split_gru_state(sd, n_layers, bidirectional)]
    hid, gsaved = gru_fwd(patches, ws, bidirectional)
    logits = linear(hid, c(sd["fc_decoder_out.weight"]), c(sd["fc_decoder_out.bias"]))
    return logits, {"front": fsaved, "gru": gsaved, "hid": hid, "ws": ws, "T": x.shape[1]}


def decoder_backward(sd, saved, dlogits, day_idx, *, kernel_len, stride_len, bidirectional, dtype=np.float64):
    """Gradients of every live parameter, keyed like the state dict."""
    c = lambda a: np.asarray(a).astype(dtype)
    B, Tp, C = dlogits.shape
    hid = saved["hid"]
    d2 = dlogits.reshape(B * Tp, C)
    g = {"fc_decoder_out.weight": d2.T @ hid.reshape(B * Tp, -1), "fc_decoder_out.bias": d2.sum(0)}
    dhid = (d2 @ c(sd["fc_decoder_out.weight"])).reshape(B, Tp, -1)
    dpatches, ggru = gru_bwd(dhid, saved["gru"], saved["ws"], bidirectional)
    for l, gl in enumerate(ggru):
        for k, v in gl.items():
            sfx = "_reverse" if k.endswith("_reverse") else ""
            base = k[:-8] if sfx else k
            name = {"w_ih": "weight_ih", "w_hh": "weight_hh", "b_ih": "bias_ih", "b_hh": "bias_hh"}[base]
            g[f"gru_decoder.{name}_l{l}{sfx}"] = v
    dW, db = frontend_bwd(dpatches, saved["front"], c(sd["dayWeights"]), day_idx, saved["T"],
                          kernel_len, stride_len)
    g["dayWeights"] = dW
    g["dayBias"] = db
    return g


def train_loss_and_grads(sd, x, day_idx, targets, x_len, y_len, *, kernel_len, stride_len, n_layers,
                         bidirectional, dtype=np.float64):
    """forward -> log_softmax -> CTC(mean, zero_infinity) -> backward
    (trainer:208-218, 242, 251-252).  Returns (loss, logits, grads)."""
    logits, saved = decoder_forward(sd, x, day_idx, kernel_len=kernel_len, stride_len=stride_len,
                                    n_layers=n_layers, bidirectional=bidirectional, dtype=dtype)
    lp = log_softmax(logits, axis=2)
    lens = out_lens(x_len, kernel_len, stride_len)
    loss, _, glp = ctc_loss(np.transpose(lp, (1, 0, 2)), targets, lens, y_len)
    glp = np.transpose(glp, (1, 0, 2)).astype(dtype)          # [B,T',C]
    # log_softmax backward: g - softmax * sum_c g
    dlogits = glp - np.exp(lp) * glp.sum(axis=2, keepdims=True)
    grads = decoder_backward(sd, saved, dlogits, day_idx, kernel_len=kernel_len, stride_len=stride_len,
                             bidirectional=bidirectional, dtype=dtype)
    return loss, logits, grads
