"""Train-mode encoder: forward with batch-statistics BatchNorm + DropPath and the hand-written backward.

Schedule of the SUN-M meta-tuning step's device work (reference: meta_tuning_sun_m/train_meta.py:168-174 driving
autograd through test_phase/models/visformer.py).  Every device operation is a libsunb200 entry point
(include/sunb200.h): sunb_gemm (forward convs, dgrads with the activation derivative fused into the epilogue),
sunb_wgrad (tcgen05 weight gradients), sunb_attention(+_backward) on the padded head layout (tcgen05 forward and backward),
sunb_pack_weights (all bf16 operand layouts of a step in one launch) and the BatchNorm / stem-tail / helper kernels.
torch is used to allocate buffers and for views/permutes of small gradient tensors.  Weight gradients run on a second stream
(on_wgrad_stream) next to the dgrad / BatchNorm chain; every kernel is launched with programmatic dependent launch.

Semantics kept from the reference: BN statistics over the whole concatenated batch, biased variance
for normalisation / unbiased for the running stats with momentum 0.1, DropPath as a per-sample scale mask/keep drawn
with torch.rand in forward order (attention branch first, then MLP), identity DropPath for rate 0.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import native as N

ACT_NONE, ACT_LRELU, ACT_GELU = 0, 1, 2
HEADS = 6
DEPTH = (4, 2, 3)
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _st():
    return N.current_stream()


def gemm(A, Wt, M, Nn, K, *, lda=None, ldw=None, out=None, ldc=None, taps=1, groups=1, a_goff=0, c_goff=0, conv=None,
         bias=None, act=ACT_NONE, resid=None, row_scale=None, rows_per_img=1, out2=None, dact_aux=None, dact=ACT_NONE,
         out_f32=None):
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups, d.a_goff, d.c_goff = M, Nn, K, taps, groups, a_goff, c_goff
    if conv is not None:
        d.a_mode, d.H, d.W, d.bw, d.bh = 1, conv[0], conv[1], conv[2], conv[3]
    d.A, d.lda = A.data_ptr(), lda if lda is not None else A.shape[-1]
    d.Wt, d.ldw = Wt.data_ptr(), ldw if ldw is not None else Wt.shape[-1]
    d.bias, d.bias_mod, d.bias_ld = N.ptr(bias), 1, 0
    d.act = act
    if resid is not None:
        d.resid, d.ldr = resid.data_ptr(), resid.shape[-1]
    if row_scale is not None:
        d.row_scale, d.rows_per_img = row_scale.data_ptr(), rows_per_img
    else:
        d.rows_per_img = 1
    if out is not None:
        d.out, d.ldc = out.data_ptr(), ldc if ldc is not None else out.shape[-1]
    if out_f32 is not None:
        d.out_f32, d.ldc_f32 = out_f32.data_ptr(), out_f32.shape[-1]
    if out2 is not None:
        d.out2, d.ldc2 = out2.data_ptr(), out2.shape[-1]
    if dact_aux is not None:
        d.dact_aux, d.ld_aux, d.dact = dact_aux.data_ptr(), dact_aux.shape[-1], dact
    N.check(N.lib().sunb_gemm(C.byref(d), 0, _st()), "sunb_gemm")
    return out


# Weight gradients are off the backward's critical path (nothing downstream reads them before the optimizer), so the ones that
# write straight into the gradient buffer run on a SIDE stream, overlapping the dgrad / BatchNorm chain.  With small per-GPU
# batches (data-parallel shards) every kernel is latency bound and the overlap is worth ~20 % of the backward; at full batch
# the two streams share the SMs and it is neutral.  TrainEngine.backward sets / joins the stream.
_wg_stream = None
_wg_keep = []


def wgrad(dY, X, out, P, Ma, Nb, *, Ca=None, Cb=None, ldo=None, taps=1, groups=1, a_goff=0, b_goff=0, conv=None, side=False):
    if side and _wg_stream is not None:
        main = torch.cuda.current_stream()
        _wg_stream.wait_stream(main)                 # producers of dY / X (and the zero fill of `out`) are enqueued on `main`
        with torch.cuda.stream(_wg_stream):
            wgrad(dY, X, out, P, Ma, Nb, Ca=Ca, Cb=Cb, ldo=ldo, taps=taps, groups=groups, a_goff=a_goff, b_goff=b_goff, conv=conv)
        _wg_keep.append((dY, X))                     # keep the operands alive until `main` has joined the side stream
        return
    d = N.WgradDesc()
    d.P, d.Ma, d.Nb = P, Ma, Nb
    d.Ca, d.Cb = Ca if Ca is not None else dY.shape[-1], Cb if Cb is not None else X.shape[-1]
    d.groups, d.a_goff, d.b_goff, d.taps = groups, a_goff, b_goff, taps
    if conv is not None:
        d.mode, d.H, d.W, d.bw, d.bh = 1, conv[0], conv[1], conv[2], conv[3]
    d.dY, d.ldy, d.X, d.ldx = dY.data_ptr(), dY.shape[-1], X.data_ptr(), X.shape[-1]
    d.out, d.ldo, d.ksplit = out.data_ptr(), ldo if ldo is not None else Nb, 0
    N.check(N.lib().sunb_wgrad(C.byref(d), _st()), "sunb_wgrad")


def on_wgrad_stream(fn, *keep):
    """Run `fn` (weight-gradient launches plus their fold into the gradient buffer) on the side stream."""
    if _wg_stream is None:
        return fn()
    _wg_stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(_wg_stream):
        fn()
    _wg_keep.append(keep)


def head_dims(dim):
    """(d, ds): channels per head (visformer.py:170-171) and the padded head stride of the tcgen05 attention kernels."""
    d = round(dim // HEADS)
    return d, (48 if d <= 48 else 96)


def weight_pack_plan(P):
    """Pure host description of every bf16 operand copy of a training step: a list of
    (key, source parameter, dims[4], source strides[4], source offset, padded last dim, row length of the 2-D view, valid2)
    with  dst[a][b][c][d] = src.flat[off + a*s0 + b*s1 + c*s2 + d*s3]  for d < dims[3] (and c < valid2 if valid2), else 0.
    `.f` = forward operand, `.d` = data-gradient operand (transposed, taps mirrored).  tests/test_host_logic.py emulates it."""
    plan = []                                   # (key, src, dims4, strides4, off, ldd, row length of the 2-D view)

    def add(key, src, dims, strides, off=0, ldd=None, cols=None, valid2=0):
        dims, strides = tuple(dims), tuple(strides)
        while len(dims) < 4:                    # leading unit dims
            dims, strides = (1,) + dims, (0,) + strides
        plan.append((key, src, dims, strides, off, ldd or dims[3], cols or ldd or dims[3], valid2))

    for name, cin, cout in (("stem.conv2", 64, 128), ("stem.conv3", 128, 128)):
        w = P[name + ".weight"]
        add(name + ".f", w, (9, cout, cin), (1, cin * 9, 9))                          # [tap][n][c]
        add(name + ".d", w, (9, cin, cout), (-1, 9, cin * 9), off=8)                   # [tap][c][n], taps mirrored
    for i in range(DEPTH[0]):
        b = f"stage1.{i}.mlp."
        add(b + "conv1.f", P[b + "conv1.weight"], (256, 128), (128, 1))
        add(b + "conv1.d", P[b + "conv1.weight"], (128, 256), (1, 128))
        add(b + "conv3.f", P[b + "conv3.weight"], (128, 256), (256, 1))
        add(b + "conv3.d", P[b + "conv3.weight"], (256, 128), (1, 256))
        # grouped [256][32][3][3] -> [8 groups][9 taps][32 n][32 k]; the dgrad copy swaps n / k and mirrors the taps
        add(b + "conv2.f", P[b + "conv2.weight"], (8, 9, 32, 32), (32 * 288, 1, 288, 9))
        add(b + "conv2.d", P[b + "conv2.weight"], (8, 9, 32, 32), (32 * 288, -1, 9, 288), off=8)
    for stage, cin, dim, depth in (("2", 128, 256, DEPTH[1]), ("3", 256, 512, DEPTH[2])):
        w = P[f"patch_embed{stage}.proj.weight"]
        add(f"pe{stage}.f", w, (dim, 4, cin), (cin * 4, 1, 4), cols=4 * cin)          # [n][(tap,c)]
        add(f"pe{stage}.d", w, (4, cin, dim), (1, 4, cin * 4))                        # [(tap,c)][n]
        d, ds = head_dims(dim)
        inner = HEADS * d
        for i in range(depth):
            b = f"stage{stage}.{i}."
            # heads padded from d to ds channels (zero weights): the tcgen05 attention kernels' layout
            wq, wp = P[b + "attn.qkv.weight"], P[b + "attn.proj.weight"]
            add(b + "qkv.f", wq, (3 * HEADS, ds, dim), (d * dim, dim, 1), valid2=d)          # [(x,y)][z][c]
            add(b + "qkv.d", wq, (dim, 3 * HEADS, d), (1, d * dim, dim), ldd=ds, cols=3 * HEADS * ds)   # [c][(x,y)][z]
            add(b + "proj.f", wp, (dim, HEADS, d), (inner, d, 1), ldd=ds, cols=HEADS * ds)     # [n][y][z]
            add(b + "proj.d", wp, (HEADS, ds, dim), (d, 1, inner), valid2=d)                    # [y][z][n]
            add(b + "conv1.f", P[b + "mlp.conv1.weight"], (4 * dim, dim), (dim, 1))
            add(b + "conv1.d", P[b + "mlp.conv1.weight"], (dim, 4 * dim), (1, dim))
            add(b + "conv3.f", P[b + "mlp.conv3.weight"], (dim, 4 * dim), (4 * dim, 1))
            add(b + "conv3.d", P[b + "mlp.conv3.weight"], (4 * dim, dim), (1, 4 * dim))
    return plan


class BNRec:
    """Per-layer BatchNorm record: statistics of this step and the tensors the backward needs."""
    __slots__ = ("name", "C", "count", "buf", "x", "frozen")

    def __init__(self, name, Cc, count, buf, x, frozen=False):
        self.name, self.C, self.count, self.buf, self.x, self.frozen = name, Cc, count, buf, x, frozen

    # buf rows: 0 sum, 1 sq, 2 scale, 3 shift, 4 mean, 5 rstd, 6 sdz, 7 sdzx, 8 a, 9 c1, 10 c2
    def row(self, i):
        return self.buf[i]


class _ZeroArena:
    """One zero-filled fp32 buffer per pass, carved into the small accumulators the kernels need zeroed (BatchNorm sums and
    tickets, pos-embed / bias / weight-gradient scratch): ONE fill launch instead of ~40 per step."""

    def __init__(self, n_floats, device):
        self.buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.buf.numel():             # sized generously; fall back to a fresh allocation rather than fail
            return torch.zeros(*shape, dtype=torch.float32, device=self.buf.device)
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 63) // 64 * 64                 # keep every slice 256-byte aligned
        return v


class TrainEngine:
    def __init__(self):
        self.lib = N.lib()
        self.frozen_bn = frozenset()       # names of BatchNorm layers currently in eval() (utils.freeze_bn)
        self.arena = None

    # ------------------------------------------------------------------ small wrappers
    def empty(self, *shape, dtype=torch.bfloat16):
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def zeros(self, *shape):
        """Zero-filled fp32 scratch from the pass's arena (created on demand for stand-alone kernel tests)."""
        if self.arena is None:
            self.arena = _ZeroArena(64 * 1024, self.dev)
        return self.arena.take(*shape)

    def bn_forward(self, x, name, Cc, M, P, Bf, update_running=True) -> BNRec:
        """colstats + finalize for BatchNorm `name` over x [M, C] (bf16).  Running stats updated in place."""
        buf = self.zeros(11, Cc)
        rec = BNRec(name, Cc, float(M), buf, x, frozen=name in self.frozen_bn)
        if rec.frozen:      # module switched to eval() by utils.freeze_bn: running statistics, no update
            N.check(self.lib.sunb_bn_frozen(P[name + ".weight"].data_ptr(), P[name + ".bias"].data_ptr(),
                                            Bf[name + ".running_mean"].data_ptr(), Bf[name + ".running_var"].data_ptr(), BN_EPS,
                                            Cc, buf[2].data_ptr(), buf[3].data_ptr(), buf[4].data_ptr(), buf[5].data_ptr(),
                                            _st()), "sunb_bn_frozen")
            return rec
        rm, rv, nbt = Bf[name + ".running_mean"], Bf[name + ".running_var"], Bf[name + ".num_batches_tracked"]
        ticket = self.zeros(64)
        # column sums + finalize (scale / shift / running statistics) in one launch: the last block finalizes
        N.check(self.lib.sunb_bn_stats_forward(x.data_ptr(), x.shape[-1], M, Cc, buf[0].data_ptr(), buf[1].data_ptr(),
                                               ticket.data_ptr(), P[name + ".weight"].data_ptr(), P[name + ".bias"].data_ptr(),
                                               rm.data_ptr() if update_running else None,
                                               rv.data_ptr() if update_running else None,
                                               nbt.data_ptr() if update_running else None, BN_MOMENTUM, BN_EPS,
                                               buf[2].data_ptr(), buf[3].data_ptr(), buf[4].data_ptr(), buf[5].data_ptr(), _st()),
                "sunb_bn_stats_forward")
        return rec

    def bn_apply(self, x, rec: BNRec, M, act=ACT_NONE, tab=None, tab_mod=1):
        out = self.empty(M, rec.C)
        N.check(self.lib.sunb_bn_apply(x.data_ptr(), x.shape[-1], rec.row(2).data_ptr(), rec.row(3).data_ptr(), act,
                                       N.ptr(tab), tab_mod, out.data_ptr(), rec.C, M, rec.C, _st()), "sunb_bn_apply")
        return out

    def bn_backward(self, dz, rec: BNRec, M, P, G, res=None):
        """dz = gradient w.r.t. the BN output [M, C] -> gradient w.r.t. its input (+ res); accumulates dgamma / dbeta."""
        b = rec.buf
        ticket = self.zeros(64)
        N.check(self.lib.sunb_bn_stats_backward(dz.data_ptr(), dz.shape[-1], rec.x.data_ptr(), rec.x.shape[-1], M, rec.C,
                                                b[6].data_ptr(), b[7].data_ptr(), ticket.data_ptr(), rec.count, b[4].data_ptr(),
                                                b[5].data_ptr(), P[rec.name + ".weight"].data_ptr(), int(rec.frozen),
                                                b[8].data_ptr(), b[9].data_ptr(), b[10].data_ptr(),
                                                G[rec.name + ".weight"].data_ptr(), G[rec.name + ".bias"].data_ptr(), _st()),
                "sunb_bn_stats_backward")
        out = self.empty(M, rec.C)
        N.check(self.lib.sunb_bn_bwd_apply(dz.data_ptr(), dz.shape[-1], rec.x.data_ptr(), rec.x.shape[-1], b[8].data_ptr(),
                                           b[9].data_ptr(), b[10].data_ptr(), b[4].data_ptr(), N.ptr(res),
                                           res.shape[-1] if res is not None else 0, out.data_ptr(), rec.C, M, rec.C, _st()),
                "sunb_bn_bwd_apply")
        return out

    def pcast(self, src, dims, strides, off=0, ldd=None):
        """bf16 dst[a][b][c] = src.flat[off + a*sa + b*sb + c*sc]."""
        A, B, Cd = dims
        ldd = ldd or Cd
        dst = self.empty(A, B, ldd)
        N.check(self.lib.sunb_permute_cast(src.data_ptr(), off, strides[0], strides[1], strides[2], A, B, Cd, ldd,
                                           dst.data_ptr(), _st()), "sunb_permute_cast")
        return dst

    def scale_rows(self, g, rs, rows_per_img, M, Cc):
        if rs is None:
            return g
        out = self.empty(M, Cc)
        N.check(self.lib.sunb_scale_rows(g.data_ptr(), rs.data_ptr(), rows_per_img, out.data_ptr(), M, Cc, _st()),
                "sunb_scale_rows")
        return out

    # ------------------------------------------------------------------ weight preparation (fp32 masters -> bf16 operands)
    def prep_weights(self, P) -> Dict[str, torch.Tensor]:
        """Forward (.f) and data-gradient (.d) bf16 operand layouts of every GEMM / conv weight, written by ONE launch.
        The descriptor table and the destination buffer are built once per parameter set and reused every step."""
        key = tuple(t.data_ptr() for t in P.values())
        cached = getattr(self, "_pack", None)
        if cached is None or cached[0] != key:
            cached = (key,) + self._plan_weight_pack(P)
            self._pack = cached
        _, descs, n, W, _buf = cached
        N.check(self.lib.sunb_pack_weights(descs, n, _st()), "sunb_pack_weights")
        return W

    def _plan_weight_pack(self, P):
        plan = weight_pack_plan(P)
        sizes = [(dm[0] * dm[1] * dm[2] * ldd + 127) // 128 * 128 for _, _, dm, _, _, ldd, _, _ in plan]  # 256-byte aligned
        buf = torch.empty(sum(sizes), dtype=torch.bfloat16, device=self.dev)
        descs = (N.PackDesc * len(plan))()
        W, o = {}, 0
        for e, ((key, src, dims, strides, off, ldd, cols, valid2), sz) in enumerate(zip(plan, sizes)):
            n_el = dims[0] * dims[1] * dims[2] * ldd
            dst = buf[o:o + n_el]
            o += sz
            descs[e].src, descs[e].dst, descs[e].off, descs[e].ldd = src.data_ptr(), dst.data_ptr(), off, ldd
            descs[e].valid2 = valid2
            for j in range(4):
                descs[e].strides[j], descs[e].dims[j] = strides[j], dims[j]
            W[key] = dst.view(-1, cols)
        return descs, len(plan), W, buf

    # ------------------------------------------------------------------ forward
    def forward(self, P, Bf, x, rs: Dict[str, List[Optional[torch.Tensor]]], update_running=True):
        """P: encoder parameters (fp32 CUDA, reference names without the 'encoder.' prefix); Bf: BN buffers;
        x fp32 [B,3,80,80]; rs[block] = per-branch DropPath scales (fp32 [B]) or None.
        Returns (pooled fp32 [B,512], dense fp32 NHWC [B,5,5,512], pooled bf16, dense bf16, ctx)."""
        self.dev = x.device
        B = x.shape[0]
        lib = self.lib
        W = self.prep_weights(P)
        ctx = {"B": B, "W": W, "x": x, "rs": rs, "blocks": []}
        self.arena = _ZeroArena(160 * 1024, self.dev)              # forward: 21 BatchNorm records (11 x C) + tickets
        z64 = self.zeros(64)
        z128 = self.zeros(128)

        # ---- stem (visformer.py:220-239)
        M0 = B * 1600
        a1r, idr = self.empty(M0, 64), self.empty(M0, 128)
        N.check(lib.sunb_stem_in(x.data_ptr(), P["stem.conv1.weight"].data_ptr(), z64.data_ptr(),
                                 P["stem.downsample.0.weight"].data_ptr(), z128.data_ptr(), a1r.data_ptr(), idr.data_ptr(),
                                 B, 0, _st()), "sunb_stem_in")
        bn1 = self.bn_forward(a1r, "stem.bn1", 64, M0, P, Bf, update_running)
        a1 = self.bn_apply(a1r, bn1, M0, ACT_LRELU)
        a2r = gemm(a1, W["stem.conv2.f"], M0, 128, 64, out=self.empty(M0, 128), taps=9, conv=(40, 40, 8, 8))
        bn2 = self.bn_forward(a2r, "stem.bn2", 128, M0, P, Bf, update_running)
        a2 = self.bn_apply(a2r, bn2, M0, ACT_LRELU)
        c3r = gemm(a2, W["stem.conv3.f"], M0, 128, 128, out=self.empty(M0, 128), taps=9, conv=(40, 40, 8, 8))
        bn3 = self.bn_forward(c3r, "stem.bn3", 128, M0, P, Bf, update_running)
        bnd = self.bn_forward(idr, "stem.downsample.1", 128, M0, P, Bf, update_running)
        pos1 = P["pos_embed1"][0].permute(1, 2, 0).reshape(400, 128).contiguous()
        M1 = B * 400
        cur = self.empty(M1, 128)
        N.check(lib.sunb_stem_tail_forward(c3r.data_ptr(), idr.data_ptr(), bn3.row(2).data_ptr(), bn3.row(3).data_ptr(),
                                           bnd.row(2).data_ptr(), bnd.row(3).data_ptr(), pos1.data_ptr(), cur.data_ptr(), B,
                                           _st()), "sunb_stem_tail_forward")
        ctx["stem"] = dict(a1r=a1r, idr=idr, a1=a1, a2r=a2r, a2=a2, c3r=c3r, bn1=bn1, bn2=bn2, bn3=bn3, bnd=bnd)

        # ---- stage 1 (visformer.py:152-163, 259-263)
        for i in range(DEPTH[0]):
            name = f"stage1.{i}"
            r = (rs.get(name) or [None])[0]
            bn = self.bn_forward(cur, name + ".norm2.bn", 128, M1, P, Bf, update_running)
            xn = self.bn_apply(cur, bn, M1)
            h1, h1p = self.empty(M1, 256), self.empty(M1, 256)
            gemm(xn, W[name + ".mlp.conv1.f"], M1, 256, 128, out=h1, act=ACT_GELU, out2=h1p)
            h2, h2p = self.empty(M1, 256), self.empty(M1, 256)
            N.check(lib.sunb_gconv3x3(h1.data_ptr(), 256, W[name + ".mlp.conv2.f"].data_ptr(), h2.data_ptr(), 256,
                                      h2p.data_ptr(), 256, None, 0, B, ACT_GELU, ACT_NONE, _st()), "sunb_gconv3x3")
            nxt = gemm(h2, W[name + ".mlp.conv3.f"], M1, 128, 256, out=self.empty(M1, 128), resid=cur, row_scale=r,
                       rows_per_img=400)
            ctx["blocks"].append(dict(kind="conv", name=name, x=cur, bn=bn, xn=xn, h1=h1, h1p=h1p, h2=h2, h2p=h2p, rs=r))
            cur = nxt

        # ---- stages 2 and 3
        for stage, cin, dim, side, depth in (("2", 128, 256, 10, DEPTH[1]), ("3", 256, 512, 5, DEPTH[2])):
            S = side * side
            M = B * S
            xs = self.empty(M, 4 * cin)
            N.check(lib.sunb_s2d_reorder(cur.data_ptr(), xs.data_ptr(), B, 2 * side, 2 * side, cin, 0, _st()), "sunb_s2d_reorder")
            y = gemm(xs, W[f"pe{stage}.f"], M, dim, 4 * cin, out=self.empty(M, dim), bias=P[f"patch_embed{stage}.proj.bias"])
            bnp = self.bn_forward(y, f"patch_embed{stage}.norm.bn", dim, M, P, Bf, update_running)
            pos = P[f"pos_embed{stage}"][0].permute(1, 2, 0).reshape(S, dim).contiguous()
            cur = self.bn_apply(y, bnp, M, ACT_NONE, tab=pos, tab_mod=S)
            ctx[f"pe{stage}"] = dict(xs=xs, y=y, bn=bnp)
            d, ds = head_dims(dim)
            inner = HEADS * ds                                       # padded heads: pad channels are exact zeros everywhere
            ldi, ld3 = inner, 3 * inner
            for i in range(depth):
                name = f"stage{stage}.{i}"
                rr = rs.get(name) or [None, None]
                bnA = self.bn_forward(cur, name + ".norm1.bn", dim, M, P, Bf, update_running)
                xn1 = self.bn_apply(cur, bnA, M)
                qkv = gemm(xn1, W[name + ".qkv.f"], M, 3 * inner, dim, out=self.empty(M, ld3))
                ao = self.empty(M, ldi)
                N.check(lib.sunb_attention(qkv.data_ptr(), ao.data_ptr(), B, S, d, ds, HEADS, ld3, ldi, _st()), "sunb_attention")
                mid = gemm(ao, W[name + ".proj.f"], M, dim, inner, out=self.empty(M, dim), resid=cur, row_scale=rr[0],
                           rows_per_img=S)
                bnM = self.bn_forward(mid, name + ".norm2.bn", dim, M, P, Bf, update_running)
                xn2 = self.bn_apply(mid, bnM, M)
                hid, hidp = self.empty(M, 4 * dim), self.empty(M, 4 * dim)
                gemm(xn2, W[name + ".conv1.f"], M, 4 * dim, dim, out=hid, act=ACT_GELU, out2=hidp)
                nxt = gemm(hid, W[name + ".conv3.f"], M, dim, 4 * dim, out=self.empty(M, dim), resid=mid, row_scale=rr[1],
                           rows_per_img=S)
                ctx["blocks"].append(dict(kind="attn", name=name, x=cur, bnA=bnA, xn1=xn1, qkv=qkv, ao=ao, mid=mid, bnM=bnM,
                                          xn2=xn2, hid=hid, hidp=hidp, rs=rr, S=S, dim=dim, d=d, ds=ds, inner=inner, ldi=ldi,
                                          ld3=ld3, M=M, stage=stage))
                cur = nxt

        # ---- final BN + pool (visformer.py:455-462)
        Mf = B * 25
        bnf = self.bn_forward(cur, "norm.bn", 512, Mf, P, Bf, update_running)
        pooled = self.empty(B, 512, dtype=torch.float32)
        dense = self.empty(B, 5, 5, 512, dtype=torch.float32)
        pooled16, dense16 = self.empty(B, 512), self.empty(B, 5, 5, 512)        # bf16 copies for the linear heads
        N.check(lib.sunb_final_norm_pool(cur.data_ptr(), bnf.row(2).data_ptr(), bnf.row(3).data_ptr(), dense.data_ptr(),
                                         dense16.data_ptr(), pooled.data_ptr(), pooled16.data_ptr(), B, 25, 512, _st()),
                "sunb_final_norm_pool")
        ctx["final"] = dict(x=cur, bn=bnf)
        return pooled, dense, pooled16, dense16, ctx

    # ------------------------------------------------------------------ backward
    # parameter groups in the order their gradients become final during the backward pass
    @staticmethod
    def _grad_groups(names):
        def pick(*prefixes):
            return [n for n in names if n.startswith(prefixes)]
        return [pick("stage3.", "norm."), pick("patch_embed3.", "pos_embed3"), pick("stage2."),
                pick("patch_embed2.", "pos_embed2"), pick("stage1."), pick("stem.", "pos_embed1")]

    def backward(self, P, ctx, dpooled, ddense=None, comm=None) -> Dict[str, torch.Tensor]:
        """Gradients of every encoder parameter.  With `comm` (a GradComm, multi-GPU data parallel) the gradients live in
        one flat buffer ordered by readiness and each group is all-reduced on a side stream as soon as its stage of
        the backward pass has been enqueued, overlapping NCCL with the remaining dgrad / wgrad kernels."""
        lib = self.lib
        B, W = ctx["B"], ctx["W"]
        groups = self._grad_groups(list(P.keys()))
        order = [n for grp in groups for n in grp]
        assert len(order) == len(P), "parameter grouping must cover every encoder parameter"
        flat = torch.zeros(sum(P[n].numel() for n in order), dtype=torch.float32, device=self.dev)
        self.arena = _ZeroArena(7168 * 1024, self.dev)              # backward: tickets, pos / bias sums, weight-gradient scratch
        G, spans, off = {}, [], 0
        for grp in groups:
            lo = off
            for n in grp:
                G[n] = flat[off:off + P[n].numel()].view_as(P[n])
                off += P[n].numel()
            spans.append((lo, off))
        done = iter(spans)

        global _wg_stream
        if getattr(self, "_side", None) is None or self._side_dev != self.dev:
            self._side, self._side_dev = torch.cuda.Stream(device=self.dev), self.dev
        _wg_stream = self._side

        def group_ready():
            lo, hi = next(done)
            if comm is not None:
                torch.cuda.current_stream().wait_stream(self._side)      # this group's side-stream weight gradients are final
                _wg_keep.clear()
                comm.all_reduce_async(flat[lo:hi])

        Mf = B * 25
        dy = self.empty(Mf, 512)
        N.check(lib.sunb_pool_backward(N.ptr(dpooled), N.ptr(ddense), dy.data_ptr(), B, 25, 512, _st()), "sunb_pool_backward")
        g = self.bn_backward(dy, ctx["final"]["bn"], Mf, P, G)

        blocks = ctx["blocks"]
        idx = len(blocks) - 1
        for stage, cin, dim, side, depth in (("3", 256, 512, 5, DEPTH[2]), ("2", 128, 256, 10, DEPTH[1])):
            for _ in range(depth):
                g = self._attn_block_backward(blocks[idx], g, P, G, W)
                idx -= 1
            group_ready()                                   # stage blocks (+ final norm for stage 3)
            # PatchEmbed + pos_embed (visformer.py:438-441, 447-450)
            S, M = side * side, B * side * side
            pe = ctx[f"pe{stage}"]
            gpos = self.zeros(S * dim)
            N.check(lib.sunb_batch_sum(g.data_ptr(), B, S * dim, gpos.data_ptr(), _st()), "sunb_batch_sum")
            G[f"pos_embed{stage}"] += gpos.view(side, side, dim).permute(2, 0, 1).unsqueeze(0)
            dyp = self.bn_backward(g, pe["bn"], M, P, G)
            sb = self.zeros(2, dim)
            N.check(lib.sunb_colstats(dyp.data_ptr(), dim, None, 0, M, dim, sb[0].data_ptr(), sb[1].data_ptr(), _st()), "colstats")
            G[f"patch_embed{stage}.proj.bias"] += sb[0]
            gw = self.zeros(dim, 4 * cin)

            def pe_wgrad(dyp=dyp, pe=pe, gw=gw, M=M, dim=dim, cin=cin, stage=stage):
                wgrad(dyp, pe["xs"], gw, M, dim, 4 * cin)
                G[f"patch_embed{stage}.proj.weight"] += gw.view(dim, 2, 2, cin).permute(0, 3, 1, 2)
            on_wgrad_stream(pe_wgrad, dyp)
            dxs = gemm(dyp, W[f"pe{stage}.d"], M, 4 * cin, dim, out=self.empty(M, 4 * cin))
            g = self.empty(M * 4, cin)
            N.check(lib.sunb_s2d_reorder(dxs.data_ptr(), g.data_ptr(), B, 2 * side, 2 * side, cin, 1, _st()), "sunb_s2d_reorder")
            group_ready()                                   # patch embed + its pos_embed
        for _ in range(DEPTH[0]):
            g = self._conv_block_backward(blocks[idx], g, P, G, W, B)
            idx -= 1
        group_ready()
        self._stem_backward(ctx, g, P, G, W, B)
        group_ready()
        torch.cuda.current_stream().wait_stream(self._side)              # join the weight-gradient stream
        _wg_keep.clear()
        _wg_stream = None
        if comm is not None:
            comm.finish(flat)
        return G

    def _attn_block_backward(self, b, g, P, G, W, ):
        lib = self.lib
        name, M, dim, S, d, inner, ldi, ld3 = b["name"], b["M"], b["dim"], b["S"], b["d"], b["inner"], b["ldi"], b["ld3"]
        ds = b["ds"]
        Bn = M // S
        r_att, r_mlp = b["rs"]
        # ---- MLP branch: out = mid + rs * conv3(gelu(conv1(BN2(mid))))
        gs = self.scale_rows(g, r_mlp, S, M, dim)
        dhp = gemm(g, W[name + ".conv3.d"], M, 4 * dim, dim, out=self.empty(M, 4 * dim), row_scale=r_mlp, rows_per_img=S,
                   dact_aux=b["hidp"], dact=ACT_GELU)
        wgrad(gs, b["hid"], G[name + ".mlp.conv3.weight"], M, dim, 4 * dim, side=True)
        dxn2 = gemm(dhp, W[name + ".conv1.d"], M, dim, 4 * dim, out=self.empty(M, dim))
        wgrad(dhp, b["xn2"], G[name + ".mlp.conv1.weight"], M, 4 * dim, dim, side=True)
        g1 = self.bn_backward(dxn2, b["bnM"], M, P, G, res=g)
        # ---- attention branch: mid = x + rs * proj(attn(qkv(BN1(x))))
        gs1 = self.scale_rows(g1, r_att, S, M, dim)
        dao = gemm(g1, W[name + ".proj.d"], M, inner, dim, out=self.empty(M, ldi), row_scale=r_att, rows_per_img=S)
        gwp = self.zeros(dim, inner)                                   # padded-head layouts; folded into the packed gradient

        def proj_wgrad():
            wgrad(gs1, b["ao"], gwp, M, dim, inner, Cb=inner)
            G[name + ".attn.proj.weight"].view(dim, HEADS, d).add_(gwp.view(dim, HEADS, ds)[:, :, :d])
        on_wgrad_stream(proj_wgrad, gs1)
        dqkv = self.empty(M, ld3)
        N.check(lib.sunb_attention_backward(b["qkv"].data_ptr(), dao.data_ptr(), dqkv.data_ptr(), Bn, S, d, ds, HEADS, ld3,
                                            ldi, _st()), "sunb_attention_backward")
        dxn1 = gemm(dqkv, W[name + ".qkv.d"], M, dim, 3 * inner, out=self.empty(M, dim))
        gwq = self.zeros(3 * inner, dim)

        def qkv_wgrad():
            wgrad(dqkv, b["xn1"], gwq, M, 3 * inner, dim, Ca=3 * inner)
            G[name + ".attn.qkv.weight"].view(3 * HEADS, d, dim).add_(gwq.view(3 * HEADS, ds, dim)[:, :d])
        on_wgrad_stream(qkv_wgrad, dqkv)
        return self.bn_backward(dxn1, b["bnA"], M, P, G, res=g1)

    def _conv_block_backward(self, b, g, P, G, W, B):
        lib = self.lib
        name, r = b["name"], b["rs"]
        M = B * 400
        gs = self.scale_rows(g, r, 400, M, 128)
        dh2p = gemm(g, W[name + ".mlp.conv3.d"], M, 256, 128, out=self.empty(M, 256), row_scale=r, rows_per_img=400,
                    dact_aux=b["h2p"], dact=ACT_GELU)
        wgrad(gs, b["h2"], G[name + ".mlp.conv3.weight"], M, 128, 256, side=True)
        dh1p = self.empty(M, 256)
        N.check(lib.sunb_gconv3x3(dh2p.data_ptr(), 256, W[name + ".mlp.conv2.d"].data_ptr(), dh1p.data_ptr(), 256, None, 0,
                                  b["h1p"].data_ptr(), 256, B, ACT_NONE, ACT_GELU, _st()), "sunb_gconv3x3(dgrad)")
        scratch = self.zeros(2 * 9 * 128, 128)

        def conv2_wgrad():
            wgrad(dh2p, b["h1"], scratch, M, 128, 128, Ca=256, Cb=256, taps=9, groups=2, a_goff=128, b_goff=128,
                  conv=(20, 20, 4, 4))
            N.check(lib.sunb_grouped_wgrad_extract(scratch.data_ptr(), G[name + ".mlp.conv2.weight"].data_ptr(), _st()),
                    "sunb_grouped_wgrad_extract")
        on_wgrad_stream(conv2_wgrad, dh2p)
        dxn = gemm(dh1p, W[name + ".mlp.conv1.d"], M, 128, 256, out=self.empty(M, 128))
        wgrad(dh1p, b["xn"], G[name + ".mlp.conv1.weight"], M, 256, 128, side=True)
        return self.bn_backward(dxn, b["bn"], M, P, G, res=g)

    def _stem_backward(self, ctx, g, P, G, W, B):
        lib = self.lib
        s = ctx["stem"]
        M0 = B * 1600
        gpos = self.zeros(400 * 128)
        N.check(lib.sunb_batch_sum(g.data_ptr(), B, 400 * 128, gpos.data_ptr(), _st()), "sunb_batch_sum")
        G["pos_embed1"] += gpos.view(20, 20, 128).permute(2, 0, 1).unsqueeze(0)
        bn3, bnd, bn2, bn1 = s["bn3"], s["bnd"], s["bn2"], s["bn1"]
        dz = self.empty(M0, 128)
        N.check(lib.sunb_stem_tail_backward(s["c3r"].data_ptr(), s["idr"].data_ptr(), bn3.row(2).data_ptr(),
                                            bn3.row(3).data_ptr(), bnd.row(2).data_ptr(), bnd.row(3).data_ptr(), g.data_ptr(),
                                            dz.data_ptr(), B, _st()), "sunb_stem_tail_backward")
        dc3r = self.bn_backward(dz, bn3, M0, P, G)
        didr = self.bn_backward(dz, bnd, M0, P, G)
        gw3 = self.zeros(9 * 128, 128)

        def conv3_wgrad():
            wgrad(dc3r, s["a2"], gw3, M0, 128, 128, taps=9, conv=(40, 40, 8, 8))
            G["stem.conv3.weight"] += gw3.view(9, 128, 128).permute(1, 2, 0).reshape(128, 128, 3, 3)
        on_wgrad_stream(conv3_wgrad, dc3r)
        da2 = gemm(dc3r, W["stem.conv3.d"], M0, 128, 128, out=self.empty(M0, 128), taps=9, conv=(40, 40, 8, 8),
                   dact_aux=s["a2"], dact=ACT_LRELU)
        da2r = self.bn_backward(da2, bn2, M0, P, G)
        gw2 = self.zeros(9 * 128, 64)

        def conv2_wgrad():
            wgrad(da2r, s["a1"], gw2, M0, 128, 64, taps=9, conv=(40, 40, 8, 8))
            G["stem.conv2.weight"] += gw2.view(9, 128, 64).permute(1, 2, 0).reshape(128, 64, 3, 3)
        on_wgrad_stream(conv2_wgrad, da2r)
        da1 = gemm(da2r, W["stem.conv2.d"], M0, 64, 128, out=self.empty(M0, 64), taps=9, conv=(40, 40, 8, 8),
                   dact_aux=s["a1"], dact=ACT_LRELU)
        da1r = self.bn_backward(da1, bn1, M0, P, G)
        patches = self.empty(M0, 32)                                  # bf16 im2col scratch of the input images
        N.check(lib.sunb_stem_wgrad(ctx["x"].data_ptr(), da1r.data_ptr(), didr.data_ptr(),
                                    G["stem.conv1.weight"].data_ptr(), G["stem.downsample.0.weight"].data_ptr(), B,
                                    patches.data_ptr(), _st()),
                "sunb_stem_wgrad")
