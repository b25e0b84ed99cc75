"""GPU bring-up diagnostics (not a test): runs the layer cases and prints mismatch patterns, then times
the big shrinker conv and the library int8 GEMM.  Usage (under gpurun): python tools/bringup.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import int_oracle  # noqa: E402
from quantv2x_b200 import _lib  # noqa: E402
from quantv2x_b200.engine import QLayer  # noqa: E402
from tests.layer_cases import make_conv, make_deconv, make_input  # noqa: E402
from tests.test_layer_gpu import CONV_CASES, DECONV_CASES  # noqa: E402

dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), "check", _lib.lib().qv2x_device_check(0), flush=True)


def report(tag, got, ref):
    got = got.astype(np.int64)
    ref = ref.astype(np.int64)
    bad = got != ref
    print(f"  {tag}: mismatch {bad.mean():.6f} ({bad.sum()} of {bad.size})", flush=True)
    if bad.any():
        idx = np.argwhere(bad)
        print("    first idx", idx[:6].tolist())
        print("    got", got[bad][:8].tolist(), "ref", ref[bad][:8].tolist())
        for ax in range(bad.ndim):
            other = tuple(a for a in range(bad.ndim) if a != ax)
            frac = bad.mean(axis=other)
            nz = np.nonzero(frac)[0]
            print(f"    axis{ax}: {len(nz)}/{bad.shape[ax]} indices bad; first {nz[:12].tolist()} frac {np.round(frac[nz[:6]], 3).tolist()}")
    return not bad.any()


ok_all = True
for case in CONV_CASES:
    name, n, H, W, cin, cout, k, stride, pad, w_bits, groups = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    p = make_conv(rng, cin, cout, k, w_bits, groups)
    x = make_input(rng, n, H, W, cin)
    acc_ref, q_ref = int_oracle.conv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"],
                                            p["out_delta"], 0.0, stride=stride, pad=pad, relu=True)
    print("conv", name, flush=True)
    try:
        layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=k,
                       stride=stride, pad=pad, w_bits=w_bits, relu=True, in_delta=p["in_delta"],
                       out_delta=p["out_delta"])
        xd = torch.from_numpy(x).to(dev)
        ho, wo = layer.out_shape(H, W)
        acc = torch.zeros((groups, n * ho * wo, cout), dtype=torch.int32, device=dev)
        y = layer.forward(xd, acc_dump=acc)
        torch.cuda.synchronize()
        a = acc.cpu().numpy().reshape(groups, n, ho, wo, cout)
        ok = report("acc", a, acc_ref)
        ok &= report("q", y.cpu().numpy(), q_ref)
        ok_all &= ok
    except Exception as e:  # noqa: BLE001
        print("  EXCEPTION", repr(e), flush=True)
        ok_all = False
        break

for case in DECONV_CASES:
    name, n, H, W, cin, cout, s = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    p = make_deconv(rng, cin, cout, s)
    x = make_input(rng, n, H, W, cin)
    acc_ref, q_ref = int_oracle.deconv_oracle(x, p["w_int"], p["w_delta"], p["w_zp"], p["bias"], p["in_delta"][0],
                                              p["out_delta"], 0.0, stride=s, relu=True)
    print("deconv", name, flush=True)
    try:
        layer = QLayer(kind=1, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=s,
                       stride=s, pad=0, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
        xd = torch.from_numpy(x).to(dev)
        acc = torch.zeros((3, n * H * W, s * s * cout), dtype=torch.int32, device=dev)
        y = layer.forward(xd, acc_dump=acc)
        torch.cuda.synchronize()
        ok = report("acc", acc.cpu().numpy(), acc_ref)
        ok &= report("q", y.cpu().numpy(), q_ref)
        ok_all &= ok
    except Exception as e:  # noqa: BLE001
        print("  EXCEPTION", repr(e), flush=True)
        ok_all = False
        break

print("ALL OK" if ok_all else "SOME FAILED", flush=True)


def time_layer(tag, n, H, W, cin, cout, groups=1, iters=20):
    rng = np.random.default_rng(1)
    p = make_conv(rng, cin, cout, 3, 8, groups)
    x = torch.from_numpy(make_input(rng, n, H, W, cin)).to(dev)
    layer = QLayer(kind=0, w_int=p["w_int"], w_delta=p["w_delta"], w_zp=p["w_zp"], bias=p["bias"], ksize=3, stride=1,
                   pad=1, w_bits=8, relu=True, in_delta=p["in_delta"], out_delta=p["out_delta"])
    cg = cin // groups
    from quantv2x_b200.engine import rowsum_u8
    rs = [rowsum_u8(x, i * cg, cg) for i in range(groups)]
    out = torch.empty((n, H, W, cout), dtype=torch.uint8, device=dev)
    for _ in range(3):
        layer.forward(x, rowsum_in=rs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        layer.forward(x, rowsum_in=rs, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ops = 2.0 * n * H * W * cout * cin * 9
    print(f"time {tag}: {ms * 1e3:.1f} us  {ops / ms / 1e9:.1f} TOP/s", flush=True)


try:
    time_layer("shrink1 256->256 100x352 n=4", 4, 100, 352, 256, 256)
    time_layer("shrink0 384->256 100x352 n=4", 4, 100, 352, 384, 256, groups=3)
    time_layer("s0 64->64 100x352 n=4", 4, 100, 352, 64, 64)
    time_layer("s1 128->128 50x176 n=4", 4, 50, 176, 128, 128)
    time_layer("s2 256->256 25x88 n=4", 4, 25, 88, 256, 256)
except Exception as e:  # noqa: BLE001
    print("timing EXCEPTION", repr(e), flush=True)

# library int8 GEMM throughput (the measured int8 tensor-pipe denominator)
try:
    M = 8192
    a = torch.randint(-100, 100, (M, M), dtype=torch.int8, device=dev)
    b = torch.randint(-100, 100, (M, M), dtype=torch.int8, device=dev)
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"int8 _int_mm 8192^3 burst: {2 * M ** 3 / best / 1e9:.1f} TOP/s", flush=True)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cnt = 0
    while time.time() - t0 < 4.0:
        for _ in range(20):
            torch._int_mm(a, b)
        cnt += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    print(f"int8 _int_mm 8192^3 sustained: {2 * M ** 3 * cnt / e0.elapsed_time(e1) / 1e9:.1f} TOP/s", flush=True)
except Exception as e:  # noqa: BLE001
    print("int_mm EXCEPTION", repr(e), flush=True)
