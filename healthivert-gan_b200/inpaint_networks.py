"""Drop-in mirror of the reference's ``models/inpaint_networks.py`` on hand-written sm_100a kernels.

Same class names, constructor arguments, attribute names and ``state_dict`` keys
(``<net>.<layer>.conv.{weight_orig,weight_u,weight_v,bias}``, ``<net>.fc_height.*``) as the
reference, so ``latest_net_G.pth`` loads unchanged and callers such as
``eval_3d_sagittal_twostage.py:100-101`` / ``pix2pix_model.py:188-189`` keep working.
All arithmetic runs in libhv_b200.so (C ABI in include/hv_b200.h); torch is only used for
device memory, parameter registration and streams.

* ``Generator.forward``      -> one native plan (hv_generator_*), reference :28-32
* ``Conv2dBlock.forward``    -> hv_sn_prepare + hv_conv2d_fwd, reference :494-503
* ``ContextualAttention``    -> hv_ctx_attn_fwd, reference :247-410
"""
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import HV_ACT, HV_SRC_DIRECT, HV_SRC_SCALAR, HV_SRC_SUB2, HV_SRC_UP2, check, ptr

_ACT_NAMES = ("relu", "elu", "sigmoid", "none", "lrelu")


class SNConv2d(nn.Module):
    """Parameter holder with the exact key layout of ``spectral_norm(nn.Conv2d(...))``:
    ``weight_orig`` / ``bias`` parameters and ``weight_u`` / ``weight_v`` buffers.  The
    initialisation draws from the torch RNG in the same order as the reference
    (Conv2d.reset_parameters, then spectral_norm's u, v), so equal seeds give equal weights."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = kernel_size, stride, padding, dilation
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        fan_in = in_channels * kernel_size * kernel_size
        bound = 1 / math.sqrt(fan_in)
        b = torch.empty(out_channels)
        nn.init.uniform_(b, -bound, bound)
        u = nn.functional.normalize(w.new_empty(out_channels).normal_(0, 1), dim=0, eps=1e-12)
        v = nn.functional.normalize(w.new_empty(fan_in).normal_(0, 1), dim=0, eps=1e-12)
        self.bias = nn.Parameter(b)
        self.weight_orig = nn.Parameter(w)
        self.register_buffer("weight_u", u)
        self.register_buffer("weight_v", v)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, spectral_norm")

    @torch.no_grad()
    def effective_weight(self, training):
        """(W_orig / sigma, sigma) on the device; runs one power iteration in place on
        ``weight_u`` / ``weight_v`` first when ``training`` (spectral_norm.py:92-114)."""
        ready = getattr(self, "_prepared", None)
        self._prepared = None
        if ready is not None and training:   # SpectralNormBatch.prepare() ran this layer's power iteration with all the others
            return ready
        w = self.weight_orig.detach().contiguous()
        w_eff = torch.empty_like(w)
        sigma = torch.empty(1, device=w.device, dtype=torch.float32)
        check(_lib.lib().hv_sn_prepare(ptr(w), ptr(self.weight_u), ptr(self.weight_v), self.out_channels,
                                       w[0].numel(), int(training), ptr(w_eff), ptr(sigma), _lib.stream()))
        if training:   # the power iteration wrote u / v through raw pointers: bump their version counters like an in-place torch op
            torch.autograd.graph.increment_version(self.weight_u)
            torch.autograd.graph.increment_version(self.weight_v)
        return w_eff, sigma


class _SpectralNormStep:
    """Buffers and job tables of ONE training forward (see SpectralNormBatch)."""
    __slots__ = ("convs", "flat", "w_eff", "dw_eff", "dw", "sigma", "tables", "pending")

    def finish(self):
        """After the tape's backward: dw_orig of every layer whose weight gradient was written (all of them, normally) in one launch."""
        from . import train_ops as T
        L = _lib.lib()
        n = len(self.convs)
        if len(self.pending) == n:
            check(L.hv_sn_bwd_multi(ptr(self.tables[6 * n:]), n, _lib.stream()))
        else:
            for i in sorted(self.pending):
                c = self.convs[i]
                check(L.hv_sn_bwd(ptr(self.dw_eff[i]), ptr(self.w_eff[i]), ptr(c.weight_u), ptr(c.weight_v), ptr(self.sigma[i:i + 1]),
                                  ptr(self.dw[i]), c.out_channels, self.w_eff[i][0].numel(), _lib.stream()))
        for i in sorted(self.pending):
            T.accumulate_param(self.convs[i].weight_orig, self.dw[i])
        self.pending = set()


class SpectralNormBatch:
    """Spectral norm of ALL conv blocks of a generator in one launch per direction (training path): the per-layer power iteration and
    adjoint are single-CTA kernels of ~25 us each, 2 x 47 of them per step (hv_sn_prepare_multi / hv_sn_bwd_multi).  Per forward one
    flat buffer holds every layer's w_eff, dw_eff (the conv weight gradient is written straight into it) and dw_orig; the device job
    tables are cached by that buffer's address (the caching allocator hands the same block back once the step has settled, so a
    steady-state step uploads nothing)."""

    def __init__(self, convs):
        self.convs = list(convs)
        self.cache = {}          # flat.data_ptr() -> (rows, device table)

    def prepare(self, tape):
        L = _lib.lib()
        convs = self.convs
        dev = convs[0].weight_orig.device
        sizes = [c.weight_orig.numel() for c in convs]
        total, n = sum(sizes), len(convs)
        st = _SpectralNormStep()
        st.convs = convs
        st.flat = flat = torch.empty(3 * total + n, device=dev, dtype=torch.float32)
        st.sigma = flat[3 * total:]
        st.w_eff, st.dw_eff, st.dw, st.pending = [], [], [], set()
        prep, bwd, off = [], [], 0
        for i, (c, sz) in enumerate(zip(convs, sizes)):
            shape = c.weight_orig.shape
            we, de, dw = (flat[k * total + off:k * total + off + sz].view(shape) for k in range(3))
            st.w_eff.append(we); st.dw_eff.append(de); st.dw.append(dw)
            dims = c.out_channels | ((sz // c.out_channels) << 32)
            sg = st.sigma[i:i + 1].data_ptr()
            u, v = c.weight_u.data_ptr(), c.weight_v.data_ptr()
            prep.append((c.weight_orig.data_ptr(), u, v, dims, we.data_ptr(), sg))
            bwd.append((de.data_ptr(), we.data_ptr(), u, v, sg, dw.data_ptr(), dims))
            off += sz
        rows = (prep, bwd)
        hit = self.cache.get(flat.data_ptr())
        if hit is None or hit[0] != rows or hit[1].device != dev:
            words = [w for r in prep for w in r] + [w for r in bwd for w in r]
            if len(self.cache) >= 8:
                self.cache.clear()
            hit = (rows, torch.tensor(words, dtype=torch.int64).to(dev))       # blocking upload; a settled step never gets here
            self.cache[flat.data_ptr()] = hit
        st.tables = hit[1]
        check(L.hv_sn_prepare_multi(ptr(st.tables), n, 1, _lib.stream()))
        for i, c in enumerate(convs):     # the power iteration wrote u / v through raw pointers: bump their version counters
            torch.autograd.graph.increment_version(c.weight_u)
            torch.autograd.graph.increment_version(c.weight_v)
            c._prepared = (st.w_eff[i], st.sigma[i:i + 1])
            c._sn_slot = (st, i)
        tape.finalizers.append(st.finish)


def conv2d_fused(sources, w_eff, bias, k, stride, pad, dil, act, hin, win, y2_head=False):
    """hv_conv2d_fwd on a channel-concatenation of ``sources`` = [(tensor, mode), ...]."""
    desc = _lib.hv_conv_desc()
    first = sources[0][0]
    n = first.shape[0]
    cin = 0
    keep = []
    for i, (t, mode) in enumerate(sources):
        t = t.contiguous()
        keep.append(t)
        ch = 1 if mode == HV_SRC_SCALAR else t.shape[1]
        desc.src[i].ptr = ptr(t)
        desc.src[i].channels = ch
        desc.src[i].mode = mode
        cin += ch
    cout = w_eff.shape[0]
    desc.n, desc.cin, desc.cout, desc.hin, desc.win = n, cin, cout, hin, win
    desc.k, desc.stride, desc.pad, desc.dil = k, stride, pad, dil
    desc.act = HV_ACT["heads"] if y2_head else HV_ACT[act]
    desc.nsrc = len(sources)
    eff = (k - 1) * dil + 1
    hout = (hin + 2 * pad - eff) // stride + 1
    wout = (win + 2 * pad - eff) // stride + 1
    dev = first.device
    if y2_head:
        y = torch.empty(n, 1, hout, wout, device=dev, dtype=torch.float32)
        y2 = torch.empty_like(y)
    else:
        y = torch.empty(n, cout, hout, wout, device=dev, dtype=torch.float32)
        y2 = None
    check(_lib.lib().hv_conv2d_fwd(desc, ptr(w_eff), ptr(bias), ptr(y), ptr(y2), _lib.stream()))
    return (y, y2) if y2_head else y


class Conv2dBlock(nn.Module):
    """reference models/inpaint_networks.py:420-503 (spectral-norm conv + bias + activation)."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, conv_padding=0, dilation=1,
                 weight_norm="sn", norm="none", activation="relu", pad_type="zero", transpose=False):
        super().__init__()
        self.use_bias = True
        assert pad_type == "zero" and padding == 0, "Unsupported padding type: {}".format(pad_type)
        assert norm == "none", "Unsupported normalization: {}".format(norm)
        assert weight_norm == "sn", "Unsupported normalization: {}".format(weight_norm)
        assert activation in _ACT_NAMES, "Unsupported activation: {}".format(activation)
        assert not transpose, "transpose convolution is not on the path"
        self.pad = None
        self.norm = None
        self.activation_name = activation
        self.conv = SNConv2d(input_dim, output_dim, kernel_size, stride, conv_padding, dilation)

    def forward(self, x, sources=None, extent=None):
        """``sources``/``extent`` (not in the reference) feed the conv from a fused channel
        concatenation [(tensor, hv_src_mode), ...] with virtual input extent (h, w)."""
        if torch.is_grad_enabled() and (x.requires_grad or self.conv.weight_orig.requires_grad and self.training):
            raise NotImplementedError("hv_b200: the backward of Conv2dBlock is not built yet; "
                                      "call under torch.no_grad()")
        c = self.conv
        w_eff, _ = c.effective_weight(self.training)
        srcs = sources if sources is not None else [(x, HV_SRC_DIRECT)]
        hin, win = extent if extent is not None else (x.shape[2], x.shape[3])
        return conv2d_fused(srcs, w_eff, c.bias.detach(), c.kernel_size, c.stride, c.padding, c.dilation,
                            self.activation_name, hin, win)


def _conv_run(block, tape, sources, extent, act=None):
    """Tape-aware Conv2dBlock forward: spectral-norm prepare (one power iteration in train mode), fused conv, and a
    recorded backward = act' -> wgrad/dgrad -> spectral-norm adjoint into ``weight_orig.grad`` / ``bias.grad``."""
    from . import train_ops as T
    c = block.conv
    slot = getattr(c, "_sn_slot", None) if (getattr(c, "_prepared", None) is not None and block.training) else None
    w_eff, sigma = c.effective_weight(block.training)

    def on_grad(dw_eff, db):
        if slot is not None:       # batched adjoint after the tape's backward (SpectralNormBatch.finish); dw_eff is its buffer already
            slot[0].pending.add(slot[1])
        else:
            dw = torch.empty_like(dw_eff)
            check(_lib.lib().hv_sn_bwd(ptr(dw_eff), ptr(w_eff), ptr(c.weight_u), ptr(c.weight_v), ptr(sigma), ptr(dw),
                                       c.out_channels, w_eff[0].numel(), _lib.stream()))
            T.accumulate_param(c.weight_orig, dw)
        T.accumulate_param(c.bias, db)

    return T.conv2d(tape, sources, w_eff, c.bias.detach(), c.kernel_size, c.stride, c.padding, c.dilation,
                    act or block.activation_name, extent, on_grad, dw_buffer=slot[0].dw_eff[slot[1]] if slot is not None else None)


def gen_conv(input_dim, output_dim, kernel_size=3, stride=1, padding=0, rate=1, activation="elu"):
    """reference models/inpaint_networks.py:413-417"""
    return Conv2dBlock(input_dim, output_dim, kernel_size, stride, conv_padding=padding, dilation=rate,
                       activation=activation)


def _height_head(x, fc):
    n, c, h, w = x.shape
    out = torch.empty(n, 1, device=x.device, dtype=torch.float32)
    xc, wc, bc = x.contiguous(), fc.weight.detach().contiguous(), fc.bias.detach().contiguous()
    check(_lib.lib().hv_gap_fc_sigmoid(ptr(xc), ptr(wc), ptr(bc), ptr(out), n, c, h * w, _lib.stream()))
    return out


def _ratio(slice_ratio, x):
    return slice_ratio.reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()


class CoarseGenerator(nn.Module):
    """reference models/inpaint_networks.py:36-117"""

    def __init__(self, input_dim, cnum, use_cuda):
        super().__init__()
        self.use_cuda = use_cuda
        self.conv1 = gen_conv(input_dim + 2, cnum, 5, 1, 2)
        self.conv2_downsample = gen_conv(cnum, cnum * 2, 3, 2, 1)
        self.conv3 = gen_conv(cnum * 2, cnum * 2, 3, 1, 1)
        self.conv4_downsample = gen_conv(cnum * 2, cnum * 4, 3, 2, 1)
        self.conv5 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.conv6 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.conv7_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 2, rate=2)
        self.conv8_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 4, rate=4)
        self.conv9_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 8, rate=8)
        self.conv10_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 16, rate=16)
        self.conv11 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.conv12 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.conv20 = gen_conv(cnum * 4 + 1, cnum * 4, 3, 1, 1)
        self.conv13 = gen_conv(cnum * 4, cnum * 2, 3, 1, 1)
        self.conv14 = gen_conv(cnum * 2, cnum * 2, 3, 1, 1)
        self.conv19 = gen_conv(cnum * 2 + 1, cnum * 2, 3, 1, 1)
        self.conv15 = gen_conv(cnum * 2, cnum, 3, 1, 1)
        self.conv16 = gen_conv(cnum, cnum // 2, 3, 1, 1)
        self.conv17 = gen_conv(cnum // 2, input_dim, 3, 1, 1, activation="none")
        self.conv18 = gen_conv(cnum // 2, input_dim, 3, 1, 1, activation="sigmoid")
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.fc_height = nn.Linear(cnum * 4, 1)

    def forward(self, x, mask, CAM, slice_ratio):
        """Layer-by-layer path through the stand-alone ops (Generator.forward uses the fused plan)."""
        ratio = _ratio(slice_ratio, x)
        mask = mask.to(x.device)
        h, w = x.shape[2:]
        t = self.conv1(x, [(x, HV_SRC_DIRECT), (ratio, HV_SRC_SCALAR), (mask, HV_SRC_DIRECT)])
        for name in ("conv2_downsample", "conv3", "conv4_downsample", "conv5", "conv6", "conv7_atrous",
                     "conv8_atrous", "conv9_atrous", "conv10_atrous"):
            t = getattr(self, name)(t)
        pred1_h = _height_head(t, self.fc_height)
        t = self.conv12(self.conv11(t))
        CAM = CAM.to(device=x.device, dtype=torch.float32)
        t = self.conv20(t, [(t, HV_SRC_UP2), (CAM, HV_SRC_SUB2)], (h // 2, w // 2))
        t = self.conv14(self.conv13(t))
        t = self.conv19(t, [(t, HV_SRC_UP2), (CAM, HV_SRC_DIRECT)], (h, w))
        t = self.conv16(self.conv15(t))
        x_stage1 = torch.clamp(self.conv17(t), -1.0, 1.0)
        coarse_seg_sigmoid = self.conv18(t)
        return coarse_seg_sigmoid, x_stage1, pred1_h


def _coarse_forward_tape(self, tape, x, mask, CAM, slice_ratio):
    """CoarseGenerator.forward on the tape (training path).  x, mask, CAM: tensors; returns Vars."""
    from . import train_ops as T
    ratio = _ratio(slice_ratio, x)
    h, w = x.shape[2:]
    run = lambda name, srcs, ext: _conv_run(getattr(self, name), tape, srcs, ext)
    t = run("conv1", [(x, HV_SRC_DIRECT), (ratio, HV_SRC_SCALAR), (mask, HV_SRC_DIRECT)], (h, w))
    ext = (h, w)
    for name in ("conv2_downsample", "conv3", "conv4_downsample", "conv5", "conv6", "conv7_atrous",
                 "conv8_atrous", "conv9_atrous", "conv10_atrous"):
        t = run(name, [(t, HV_SRC_DIRECT)], ext)
        ext = tuple(t.data.shape[2:])
    pred1_h = T.gap_fc_sigmoid(tape, t, self.fc_height)
    t = run("conv11", [(t, HV_SRC_DIRECT)], ext)
    t = run("conv12", [(t, HV_SRC_DIRECT)], ext)
    t = run("conv20", [(t, HV_SRC_UP2), (CAM, HV_SRC_SUB2)], (h // 2, w // 2))
    t = run("conv13", [(t, HV_SRC_DIRECT)], (h // 2, w // 2))
    t = run("conv14", [(t, HV_SRC_DIRECT)], (h // 2, w // 2))
    t = run("conv19", [(t, HV_SRC_UP2), (CAM, HV_SRC_DIRECT)], (h, w))
    t = run("conv15", [(t, HV_SRC_DIRECT)], (h, w))
    t = run("conv16", [(t, HV_SRC_DIRECT)], (h, w))
    # conv17 (activation 'none') followed by torch.clamp(-1, 1) (:115): fused as the clamp epilogue
    x_stage1 = _conv_run(self.conv17, tape, [(t, HV_SRC_DIRECT)], (h, w), act="clamp1")
    coarse_seg = run("conv18", [(t, HV_SRC_DIRECT)], (h, w))
    return coarse_seg, x_stage1, pred1_h


def _fine_forward_tape(self, tape, xin, x_stage1, mask, coarse_seg, slice_ratio):
    """FineGenerator.forward on the tape (training path); x_stage1 / coarse_seg are Vars (gradients reach the coarse net)."""
    from . import train_ops as T
    ratio = _ratio(slice_ratio, xin)
    h, w = xin.shape[2:]
    run = lambda name, srcs, ext: _conv_run(getattr(self, name), tape, srcs, ext)
    xnow = [(xin, HV_SRC_DIRECT), (coarse_seg, HV_SRC_DIRECT), (mask, HV_SRC_DIRECT), (ratio, HV_SRC_SCALAR)]
    t = run("conv1", xnow, (h, w))
    ext = (h, w)
    for name in ("conv2_downsample", "conv3", "conv4_downsample", "conv5", "conv6", "conv7_atrous",
                 "conv8_atrous", "conv9_atrous", "conv10_atrous"):
        t = run(name, [(t, HV_SRC_DIRECT)], ext)
        ext = tuple(t.data.shape[2:])
    x_hallu = t
    t = run("pmconv1", xnow, (h, w))
    ext = (h, w)
    for name in ("pmconv2_downsample", "pmconv3", "pmconv4_downsample", "pmconv5", "pmconv6"):
        t = run(name, [(t, HV_SRC_DIRECT)], ext)
        ext = tuple(t.data.shape[2:])
    ca = self.contextul_attention
    t, offset_flow, offsets = T.ctx_attention(tape, t, mask, ca.softmax_scale, ca.fuse, ca.per_sample_mask)
    ca.last_offsets = offsets
    t = run("pmconv9", [(t, HV_SRC_DIRECT)], ext)
    pm = run("pmconv10", [(t, HV_SRC_DIRECT)], ext)
    t = run("allconv11", [(x_hallu, HV_SRC_DIRECT), (pm, HV_SRC_DIRECT)], ext)
    pred2_h = T.gap_fc_sigmoid(tape, t, self.fc_height)
    t = run("allconv12", [(t, HV_SRC_DIRECT)], ext)
    t = run("allconv19", [(t, HV_SRC_DIRECT)], ext)
    t = run("allconv13", [(t, HV_SRC_UP2)], (h // 2, w // 2))
    t = run("allconv14", [(t, HV_SRC_DIRECT)], (h // 2, w // 2))
    t = run("allconv15", [(t, HV_SRC_UP2)], (h, w))
    t = run("allconv16", [(t, HV_SRC_DIRECT)], (h, w))
    cat = [(t, HV_SRC_DIRECT), (x_stage1, HV_SRC_DIRECT)]
    x_stage2 = _conv_run(self.allconv17, tape, cat, (h, w), act="clamp1")   # :230
    fine_seg = run("allconv18", cat, (h, w))
    return fine_seg, x_stage2, offset_flow, pred2_h


class FineGenerator(nn.Module):
    """reference models/inpaint_networks.py:120-232"""

    def __init__(self, input_dim, cnum, use_cuda=True):
        super().__init__()
        self.use_cuda = use_cuda
        self.conv1 = gen_conv(input_dim + 3, cnum, 5, 1, 2)
        self.conv2_downsample = gen_conv(cnum, cnum, 3, 2, 1)
        self.conv3 = gen_conv(cnum, cnum * 2, 3, 1, 1)
        self.conv4_downsample = gen_conv(cnum * 2, cnum * 2, 3, 2, 1)
        self.conv5 = gen_conv(cnum * 2, cnum * 4, 3, 1, 1)
        self.conv6 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.conv7_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 2, rate=2)
        self.conv8_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 4, rate=4)
        self.conv9_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 8, rate=8)
        self.conv10_atrous = gen_conv(cnum * 4, cnum * 4, 3, 1, 16, rate=16)
        self.pmconv1 = gen_conv(input_dim + 3, cnum, 5, 1, 2)
        self.pmconv2_downsample = gen_conv(cnum, cnum, 3, 2, 1)
        self.pmconv3 = gen_conv(cnum, cnum * 2, 3, 1, 1)
        self.pmconv4_downsample = gen_conv(cnum * 2, cnum * 4, 3, 2, 1)
        self.pmconv5 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.pmconv6 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1, activation="relu")
        self.contextul_attention = ContextualAttention(self.use_cuda, ksize=3, stride=1, rate=2, fuse_k=3,
                                                       softmax_scale=10, fuse=True)
        self.pmconv9 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.pmconv10 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.allconv11 = gen_conv(cnum * 8, cnum * 4, 3, 1, 1)
        self.allconv19 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.allconv12 = gen_conv(cnum * 4, cnum * 4, 3, 1, 1)
        self.allconv13 = gen_conv(cnum * 4, cnum * 2, 3, 1, 1)
        self.allconv14 = gen_conv(cnum * 2, cnum * 2, 3, 1, 1)
        self.allconv15 = gen_conv(cnum * 2, cnum, 3, 1, 1)
        self.allconv16 = gen_conv(cnum, cnum // 2, 3, 1, 1)
        self.allconv17 = gen_conv(cnum // 2 + 1, 1, 3, 1, 1, activation="none")
        self.allconv18 = gen_conv(cnum // 2 + 1, 1, 3, 1, 1, activation="sigmoid")
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.fc_height = nn.Linear(cnum * 4, 1)

    def forward(self, xin, x_stage1, mask, coarse_seg, slice_ratio):
        ratio = _ratio(slice_ratio, xin)
        mask = mask.to(xin.device)
        h, w = xin.shape[2:]
        xnow = [(xin, HV_SRC_DIRECT), (coarse_seg, HV_SRC_DIRECT), (mask, HV_SRC_DIRECT), (ratio, HV_SRC_SCALAR)]
        t = self.conv1(xin, xnow)
        for name in ("conv2_downsample", "conv3", "conv4_downsample", "conv5", "conv6", "conv7_atrous",
                     "conv8_atrous", "conv9_atrous", "conv10_atrous"):
            t = getattr(self, name)(t)
        x_hallu = t
        t = self.pmconv1(xin, xnow)
        for name in ("pmconv2_downsample", "pmconv3", "pmconv4_downsample", "pmconv5", "pmconv6"):
            t = getattr(self, name)(t)
        t, offset_flow = self.contextul_attention(t, t, mask)
        pm = self.pmconv10(self.pmconv9(t))
        t = self.allconv11(x_hallu, [(x_hallu, HV_SRC_DIRECT), (pm, HV_SRC_DIRECT)])
        pred2_h = _height_head(t, self.fc_height)
        t = self.allconv19(self.allconv12(t))
        t = self.allconv13(t, [(t, HV_SRC_UP2)], (h // 2, w // 2))
        t = self.allconv14(t)
        t = self.allconv15(t, [(t, HV_SRC_UP2)], (h, w))
        t = self.allconv16(t)
        cat = [(t, HV_SRC_DIRECT), (x_stage1, HV_SRC_DIRECT)]
        x_stage2 = torch.clamp(self.allconv17(t, cat), -1.0, 1.0)
        fine_seg_sigmoid = self.allconv18(t, cat)
        return fine_seg_sigmoid, x_stage2, offset_flow, pred2_h


class ContextualAttention(nn.Module):
    """reference models/inpaint_networks.py:235-410 (f is matched against itself: forward(f, f, mask))."""

    def __init__(self, use_cuda, ksize=3, stride=1, rate=1, fuse_k=3, softmax_scale=10, fuse=False):
        super().__init__()
        self.ksize, self.stride, self.rate, self.fuse_k = ksize, stride, rate, fuse_k
        self.softmax_scale, self.fuse, self.use_cuda = softmax_scale, fuse, use_cuda
        self.per_sample_mask = False  # False == the reference (mask of sample 0 for the whole batch)
        self.precision = "fp32"       # 'bf16': similarity / paste contractions on tcgen05 (hv_ctx_attn_fwd_bf16)
        self.last_offsets = None

    def forward(self, f, b, mask=None):
        if b is not f and not (b.data_ptr() == f.data_ptr() and b.shape == f.shape):
            raise NotImplementedError("hv_b200 ContextualAttention: only forward(x, x, mask) is built")
        if not (self.ksize == 3 and self.stride == 1 and self.rate == 2 and self.fuse_k == 3):
            raise NotImplementedError("hv_b200 ContextualAttention: only ksize=3, stride=1, rate=2, fuse_k=3")
        n, c, h, w = f.shape
        f = f.contiguous()
        if mask is None:
            mask = torch.zeros(n, 1, 4 * h, 4 * w, device=f.device, dtype=torch.float32)
        mask = mask.to(f.device).contiguous()
        y = torch.empty_like(f)
        offsets = torch.empty(n, 2, h // 2, w // 2, device=f.device, dtype=torch.int32)
        flow = torch.empty(n, 3, 4 * h, 4 * w, device=f.device, dtype=torch.float32)
        L = _lib.lib()
        if self.precision == "bf16":
            check(L.hv_ctx_attn_fwd_bf16(ptr(f), ptr(mask), ptr(y), ptr(offsets), ptr(flow), n, c, h, w,
                                         float(self.softmax_scale), int(bool(self.fuse)), int(self.per_sample_mask),
                                         _lib.stream()))
        else:
            ws = torch.empty(L.hv_ctx_attn_workspace_bytes(n, c, h, w), device=f.device, dtype=torch.uint8)
            check(L.hv_ctx_attn_fwd(ptr(f), ptr(mask), ptr(y), ptr(offsets), ptr(flow), n, c, h, w,
                                    float(self.softmax_scale), int(bool(self.fuse)), int(self.per_sample_mask), ptr(ws),
                                    _lib.stream()))
        self.last_offsets = offsets
        return y, flow


class Generator(nn.Module):
    """reference models/inpaint_networks.py:16-32.

    ``forward`` runs the whole two-stage network as one native plan.  Extra attributes (not
    in the reference): ``precision`` ('fp32' parity mode | 'bf16' tensor-core mode),
    ``per_sample_mask`` (attention mask per sample instead of the reference's sample-0
    quirk), ``return_flow`` (skip the colour-wheel image when False)."""

    def __init__(self, config, use_cuda):
        super().__init__()
        self.input_dim = config["input_dim"]
        self.cnum = config["ngf"]
        self.use_cuda = use_cuda
        self.coarse_generator = CoarseGenerator(self.input_dim, self.cnum, self.use_cuda)
        self.fine_generator = FineGenerator(self.input_dim, self.cnum, self.use_cuda)
        self.precision = "fp32"
        self.per_sample_mask = False
        self.return_flow = True
        self._plan = None
        self._plan_key = None
        self._param_sig = None
        self._sig_slots = None
        self.static_weights = False   # True: skip the per-forward weight-change check (weights frozen after the first forward)
        self.last_offsets = None
        self._last_out = None
        self._pipelines = []          # weak references to SlicePipelines built on the native plan (closed before the plan goes)
        self._u8_pipe = None

    # ---- native plan management -------------------------------------------------------
    def _layers(self):
        L = _lib.lib()
        out = []
        name = _lib.ctypes.create_string_buffer(64)
        for i in range(L.hv_generator_num_layers()):
            check(L.hv_generator_layer_info(i, name, None, None, None, None, None, None, None))
            net, layer = name.value.decode().split(".")
            out.append(getattr(getattr(self, net), layer).conv)
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)   # .cuda() / .to() / .float(): storages may have been swapped under the same tensors
        self._param_sig = None
        return out

    def train(self, mode=True):
        # a mode switch always re-prepares the plan: the train-mode forward power-iterates u / v inside the library (raw pointers,
        # no version bump), so the signature taken before it says nothing about the weights the next eval forward must use
        self._param_sig = None
        return super().train(mode)

    def _destroy_plan(self):
        for ref in getattr(self, "_pipelines", []):
            pipe = ref()
            if pipe is not None:
                pipe.close()
        self._pipelines = []
        self._u8_pipe = None
        if self._plan is not None:
            _lib.lib().hv_generator_destroy(self._plan)
            self._plan = None
            self._plan_key = None
            self._param_sig = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def _ensure_plan(self, n, device):
        if self.input_dim != 1 or self.cnum != 16:
            raise NotImplementedError("hv_b200 Generator plan is built for input_dim=1, ngf=16 (the reference's "
                                      "hard-coded configuration, pix2pix_model.py:103)")
        L = _lib.lib()
        key = (device.index, self.precision)
        if self._plan is None or self._plan_key[:2] != key or self._plan_key[2] < n:
            self._destroy_plan()
            cap = max(n, 16)
            handle = _lib.c_void_p()
            check(L.hv_generator_create(_lib.ctypes.byref(handle), cap, _lib.HV_PREC[self.precision]))
            self._plan = handle
            self._plan_key = key + (cap,)
        training = self.training
        if self.static_weights and self._param_sig is not None and not training:
            return self._plan          # the caller vouches that the weights do not change between forwards (eval loops)
        # Weight signature: identity + version counter of the 192 tensors, read through the modules' own dicts (Module.__getattr__
        # and data_ptr() made this check cost as much host time as the 59 kernel launches of the forward).  In-place updates
        # (torch optimizers, load_state_dict, and this package's own FusedAdam.step / train-mode power iteration, which call
        # increment_version after writing through raw pointers) bump the version; .cuda() / .to() go through _apply and
        # train() / eval() through train(), both of which drop the signature.
        if self._sig_slots is None:
            convs = self._layers()
            fcs = (self.coarse_generator.fc_height, self.fine_generator.fc_height)
            slots = []
            for c in convs:
                slots += [(c._parameters, "weight_orig"), (c._parameters, "bias"), (c._buffers, "weight_u"), (c._buffers, "weight_v")]
            for f in fcs:
                slots += [(f._parameters, "weight"), (f._parameters, "bias")]
            self._sig_slots = (convs, fcs, slots)
        convs, fcs, slots = self._sig_slots
        sig = [(id(t), t._version) for t in [d[k] for d, k in slots]]
        if sig != self._param_sig or training:
            for i, c in enumerate(convs):
                for t in (c.weight_orig, c.bias, c.weight_u, c.weight_v):
                    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                        raise _lib.HvError("Generator parameters must be contiguous fp32 CUDA tensors "
                                           "(call .cuda() / .to(device) first)")
                check(L.hv_generator_set_layer(self._plan, i, ptr(c.weight_orig.data), ptr(c.weight_u),
                                               ptr(c.weight_v), ptr(c.bias.data)))
            for i, f in enumerate(fcs):
                check(L.hv_generator_set_fc(self._plan, i, ptr(f.weight.data), ptr(f.bias.data)))
            check(L.hv_generator_prepare(self._plan, int(training), _lib.stream()))
            if training:   # the library power-iterated u / v in place through raw pointers
                for c in convs:
                    torch.autograd.graph.increment_version(c.weight_u)
                    torch.autograd.graph.increment_version(c.weight_v)
            self._param_sig = None if training else sig
        return self._plan

    def forward(self, x, mask, CAM, slice_ratio):
        if not x.is_cuda:
            raise _lib.HvError("hv_b200 Generator needs CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("hv_b200 Generator: autograd backward is not built yet; run the forward "
                                      "under torch.no_grad() (eval driver / evaluate_model path)")
        dev = x.device
        n, _, h, w = x.shape
        if (h, w) != (256, 256):
            raise NotImplementedError("hv_b200 Generator plan is built for 256x256 slices")
        with torch.no_grad():
            x = x.to(torch.float32).contiguous()
            mask = mask.to(device=dev, dtype=torch.float32).contiguous()
            CAM = CAM.to(device=dev, dtype=torch.float32).contiguous()
            ratio = _ratio(slice_ratio, x)
            plan = self._ensure_plan(n, dev)
            mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            coarse_seg, fine_seg, x_stage1, x_stage2 = (mk(n, 1, h, w) for _ in range(4))
            flow = mk(n, 3, h, w) if self.return_flow else None
            pred1_h, pred2_h = mk(n, 1), mk(n, 1)
            offsets = torch.empty(n, 2, 32, 32, device=dev, dtype=torch.int32)
            check(_lib.lib().hv_generator_forward(
                plan, ptr(x), ptr(mask), ptr(CAM), ptr(ratio), n, ptr(coarse_seg), ptr(fine_seg), ptr(x_stage1),
                ptr(x_stage2), ptr(flow), ptr(pred1_h), ptr(pred2_h), ptr(offsets), int(self.per_sample_mask),
                _lib.stream()))
            self.last_offsets = offsets
        out = (coarse_seg, fine_seg, x_stage1, x_stage2, flow, pred1_h, pred2_h)
        self._last_out = out  # the head taps of read_tap() alias these buffers
        return out

    @torch.no_grad()
    def forward_u8(self, ct_u8, cam_u8, rows, slice_ratio):
        """The generator behind the uint8 HOST interface of the eval driver (eval_3d_sagittal_twostage.py:84-121): numpy / CPU
        arrays in, numpy arrays out, one H2D copy + one CUDA-graph launch + one D2H copy (healthivert_gan_b200.pipeline).
        ct_u8, cam_u8: [n, 256, 256] uint8 (the composed CT plane and CAM * 255; the network sees (u8/255 - 0.5)/0.5 and
        1 - u8/255); rows: [n, 2] mask row range [r0, r1); slice_ratio: [n].
        Returns (ct_u8_out = trunc((x_stage2 + 1) * 127.5), fine_mask, coarse_mask in {0,1}, pred1_h, pred2_h)."""
        import numpy as np
        from .pipeline import SlicePipeline
        ct_u8 = np.asarray(ct_u8, dtype=np.uint8)
        n = ct_u8.shape[0]
        pipe = self._u8_pipe
        if pipe is None or pipe._h is None or pipe.batch < n:
            pipe = self._u8_pipe = SlicePipeline(self, batch=max(n, 16), depth=1)
        pipe.refresh()
        s = pipe.slot(0)
        s.ct[:n] = ct_u8.reshape(n, 256, 256)
        s.cam[:n] = np.asarray(cam_u8, dtype=np.uint8).reshape(n, 256, 256)
        s.rows[:n] = np.asarray(rows, dtype=np.int32).reshape(n, 2)
        s.ratio[:n] = np.asarray(slice_ratio, dtype=np.float32).reshape(n)
        pipe.submit(0, n)
        pipe.wait(0)
        return (s.ct_out[:n].copy(), s.fine_mask[:n].copy(), s.coarse_mask[:n].copy(), s.heights[0, :n].copy(), s.heights[1, :n].copy())

    def forward_tape(self, tape, x, mask, CAM, slice_ratio):
        """Training-path forward (layer by layer, fp32) recorded on ``tape``: returns the reference's 7-tuple with tape
        Vars in place of the differentiable outputs (coarse_seg, fine_seg, x_stage1, x_stage2, flow tensor, pred1_h, pred2_h)."""
        if not x.is_cuda:
            raise _lib.HvError("hv_b200 Generator needs CUDA tensors (no CPU fallback)")
        dev = x.device
        x = x.to(torch.float32).contiguous()
        mask = mask.to(device=dev, dtype=torch.float32).contiguous()
        CAM = CAM.to(device=dev, dtype=torch.float32).contiguous()
        if self.training and tape is not None:
            if getattr(self, "_sn_batch", None) is None:
                self._sn_batch = SpectralNormBatch(m.conv for m in self.modules() if isinstance(m, Conv2dBlock))
            self._sn_batch.prepare(tape)
        coarse_seg, x_stage1, pred1_h = _coarse_forward_tape(self.coarse_generator, tape, x, mask, CAM, slice_ratio)
        fine_seg, x_stage2, flow, pred2_h = _fine_forward_tape(self.fine_generator, tape, x, x_stage1, mask, coarse_seg,
                                                              slice_ratio)
        self.last_offsets = self.fine_generator.contextul_attention.last_offsets
        return coarse_seg, fine_seg, x_stage1, x_stage2, flow, pred1_h, pred2_h

    def run_layer(self, idx, n):
        """Measurement hook: launch the tensor-core conv kernel of layer ``idx`` alone (bf16 plan)."""
        check(_lib.lib().hv_generator_run_layer(self._plan, idx, n, _lib.stream()))

    def run_chain(self, first, count, n):
        """Measurement hook: layers first .. first+count-1 back to back, each on its predecessor's output (bf16 plan)."""
        check(_lib.lib().hv_generator_run_chain(self._plan, first, count, n, _lib.stream()))

    @torch.no_grad()
    def read_tap(self, idx):
        """Activation of conv block ``idx`` (state_dict order; 47 = attention output) of the last forward."""
        L = _lib.lib()
        # the largest activation per slice is 32 channels at 256 x 256 (conv19 / the merged fine conv1 | pmconv1 layer)
        buf = torch.empty(self._plan_key[2] * 32 * 256 * 256, device=self.coarse_generator.fc_height.weight.device)
        cnt = check(L.hv_generator_read_tap(self._plan, idx, ptr(buf), _lib.stream()))
        return buf[:cnt].clone()
