"""Drop-in mirror of the hot-path parts of the reference's ``models/networks.py`` on hand-written kernels:
``NLayerDiscriminator`` (:555-602), ``GANLoss`` (:212-278, vanilla / BCE-with-logits), ``define_D`` (:163-206),
``init_weights`` / ``init_net`` (:68-117), ``get_scheduler`` ('linear', :52-56).

The torch ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects are parameter CONTAINERS only (identical ``state_dict`` keys
``model.{0,2,3,5,6,8,9,11}.*`` and identical initialisation streams); all arithmetic runs in libhv_b200.so.
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init

from . import _lib, train_ops as T
from ._lib import HV_SRC_DIRECT


def get_norm_layer(norm_type="batch"):
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    raise NotImplementedError("normalization layer [%s] is not built (the path uses 'batch')" % norm_type)


class NLayerDiscriminator(nn.Module):
    """PatchGAN discriminator (reference models/networks.py:555-602): conv4x4 s2 + LReLU, (conv4x4 s2 + BN + LReLU) x2,
    conv4x4 s1 + BN + LReLU, conv4x4 s1 -> 1-channel logits."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d):
        super().__init__()
        if type(norm_layer) == functools.partial:
            use_bias = norm_layer.func == nn.InstanceNorm2d
        else:
            use_bias = norm_layer == nn.InstanceNorm2d
        kw, padw = 4, 1
        sequence = [nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                         norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                     norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        sequence += [nn.Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw)]
        self.model = nn.Sequential(*sequence)
        # 'fp32': SIMT parity kernels everywhere; 'bf16': the BatchNorm-followed 4x4 convs (64->128, 128->256, 256->512: 97 % of the
        # discriminator's FLOPs) run on the tensor cores with bf16 operands and fp32 accumulation (train_ops.conv2d_tc)
        self.precision = "fp32"

    def run(self, x, tape=None, param_grads=True):
        """Forward through the hand-written kernels.  x: tensor or tape Var; returns a Var (logits [N,1,30,30])."""
        if not isinstance(x, T.Var):
            x = T.Var(x.to(torch.float32).contiguous(), requires_grad=False)
        if not x.data.is_cuda:
            raise _lib.HvError("hv_b200 NLayerDiscriminator needs CUDA tensors (no CPU fallback)")
        mods = list(self.model)
        i = 0
        while i < len(mods):
            m = mods[i]
            if not isinstance(m, nn.Conv2d):
                raise NotImplementedError("unexpected module in the PatchGAN sequence: %r" % (m,))
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            fused_lrelu = isinstance(nxt, nn.LeakyReLU)
            w = m.weight.detach()
            b = m.bias.detach() if m.bias is not None else None

            def on_grad(dw, db, m=m):
                T.accumulate_param(m.weight, dw)
                if db is not None and m.bias is not None:
                    T.accumulate_param(m.bias, db)

            h, wd = x.data.shape[2], x.data.shape[3]
            tc = (self.precision == "bf16" and b is None and not fused_lrelu and m.kernel_size == (4, 4) and m.padding == (1, 1)
                  and m.dilation == (1, 1) and m.stride[0] in (1, 2) and m.in_channels % 8 == 0 and m.out_channels % 128 == 0)
            if tc:
                x = T.conv2d_tc(tape, x, w.contiguous(), m.stride[0], on_grad, param_grads)
            else:
                x = T.conv2d(tape, [(x, HV_SRC_DIRECT)], w, b, m.kernel_size[0], m.stride[0], m.padding[0], m.dilation[0],
                             "lrelu" if fused_lrelu else "none", (h, wd), on_grad, param_grads)
            i += 2 if fused_lrelu else 1
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
                if not self.training:
                    raise NotImplementedError("hv_b200 NLayerDiscriminator: eval-mode BatchNorm is not on the training path")
                x = T.bn_lrelu(tape, x, mods[i], 0.2, param_grads)
                i += 2  # BatchNorm2d + LeakyReLU
        return x

    def forward(self, input):
        return self.run(input).data


class GANLoss(nn.Module):
    """reference models/networks.py:212-278; only the 'vanilla' objective (BCEWithLogitsLoss) is on the path."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        if gan_mode != "vanilla":
            raise NotImplementedError("gan mode %s not implemented" % gan_mode)

    def __call__(self, prediction, target_is_real):
        return T.bce_logits_const(prediction, target_is_real)[0]

    def grad(self, prediction, target_is_real, g=1.0):
        return T.bce_logits_const_grad(prediction.contiguous(), target_is_real, g)


def init_weights(net, init_type="normal", init_gain=0.02):
    """reference models/networks.py:68-99 (same RNG consumption order: net.apply)."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):
    """reference :102-117.  No nn.DataParallel wrapper: multi-GPU is one process per GPU, so the
    state_dict keys are the unwrapped ones that save_networks writes anyway (base_model.py:152-173)."""
    if len(gpu_ids) > 0:
        net.to(torch.device("cuda", gpu_ids[0]))
    init_weights(net, init_type, init_gain=init_gain)
    return net


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02, gpu_ids=[]):
    """reference models/networks.py:163-206 ('basic' and 'n_layers' are the PatchGAN variants)."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer)
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


class LinearLR:
    """'linear' policy of get_scheduler (reference :52-56) for FusedAdam."""

    def __init__(self, optimizer, opt):
        self.optimizer, self.opt, self.epoch = optimizer, opt, 0

    def step(self):
        self.epoch += 1
        o = self.opt
        f = 1.0 - max(0, self.epoch + o.epoch_count - o.n_epochs) / float(o.n_epochs_decay + 1)
        for g in self.optimizer.param_groups:
            g["lr"] = g["initial_lr"] * f


def get_scheduler(optimizer, opt):
    if opt.lr_policy == "linear":
        return LinearLR(optimizer, opt)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)
