#!/usr/bin/env python
"""Stand-alone check of hv_conv2d_wgrad_bf16 / hv_conv2d_dgrad_bf16 on one geometry against fp64 autograd on bf16-rounded operands.
usage: python tools/dbg_wgrad.py n cin cout k stride pad dil h w"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from healthivert_gan_b200 import train_ops as T

n, cin, cout, k, stride, pad, dil, h, w = (int(a) for a in sys.argv[1:10])
T.BACKWARD_PRECISION = "bf16"
bf = lambda t: t.to(torch.bfloat16).to(torch.float64)
g = torch.Generator().manual_seed(1)
x = torch.randn(n, cin, h, w, generator=g)
wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
b = torch.randn(cout, generator=g)
tape = T.Tape()
xv = T.Var(x.cuda())
got = {}
out = T.conv2d(tape, [(xv, 0)], wt.cuda(), b.cuda(), k, stride, pad, dil, "none", (h, w), lambda dw, db: got.update(dw=dw, db=db))
dy = torch.randn(out.data.shape, generator=g)
out.grad = dy.cuda()
tape.backward()
torch.cuda.synchronize()
xb, wb = bf(x).requires_grad_(), bf(wt).requires_grad_()
F.conv2d(xb, wb, None, stride=stride, padding=pad, dilation=dil).backward(bf(dy))
rel = lambda a, r: float((a.double().cpu() - r).norm() / r.norm())
ex = (xv.grad.double().cpu() - xb.grad).abs()
print(sys.argv[1:], "dw rel", rel(got["dw"], wb.grad), "dx rel", rel(xv.grad, xb.grad), "dx max err", float(ex.max()), "at", [int(v) for v in torch.unravel_index(ex.argmax(), ex.shape)],
      "dx max", float(xb.grad.abs().max()))
sg = (xv.grad.double().cpu() - xb.grad)
print("   dx signed mean err / mean |dx|", float(sg.mean() / xb.grad.abs().mean()), " err correlated with sign(dx):", float((sg * xb.grad.sign()).mean() / xb.grad.abs().mean()),
      " sum(dx) got / ref", float(xv.grad.double().sum()), float(xb.grad.sum()))
# where do the errors sit?  mean abs error per image row / column
print("   dx err by row (first/last 3, mean)", [round(float(v), 5) for v in ex.mean(dim=(0, 1, 3))[:3]], [round(float(v), 5) for v in ex.mean(dim=(0, 1, 3))[-3:]], round(float(ex.mean()), 5))
print("   dx err by col (first/last 3)", [round(float(v), 5) for v in ex.mean(dim=(0, 1, 2))[:3]], [round(float(v), 5) for v in ex.mean(dim=(0, 1, 2))[-3:]])
