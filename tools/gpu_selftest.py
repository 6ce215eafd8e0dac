"""Battery of on-GPU kernel checks with verbose diagnostics (development tool; the judged tests are tests/).
Each group runs in its own subprocess under a timeout so that a trapping kernel cannot take the rest down.

    python tools/gpu_selftest.py            # all groups
    python tools/gpu_selftest.py tc_basic   # one group, in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GROUPS = ["simt", "tc_basic", "tc_shapes", "tc_stride2", "tc_epi", "vq", "accuracy", "model_simt", "model_tc", "timing"]


def _case(eng, tag, **kw):
    import torch
    from convcase import compare, run_case
    want = kw.pop("want", ("f32",))
    try:
        out, exp, _ = run_case(eng, want=want, **kw)
        err = compare(out, exp, want, kw.get("passes", 3))
        flag = eng.lib.mcq_device_error_flag()
        print(f"  {tag:44s} rel.err {err} flag={flag}", flush=True)
        return err
    except Exception as e:  # noqa
        print(f"  {tag:44s} EXCEPTION {type(e).__name__}: {e}", flush=True)
        raise


def g_simt():
    from mcquic_b200 import _lib
    from mcquic_b200.engine import Engine
    eng = Engine("simt")
    _case(eng, "3x3 s1 c64->64 12x20 n2", n=2, h=12, w=20, cin=64, cout=64)
    _case(eng, "3x3 s1 c32->40 (cin%64!=0)", n=1, h=9, w=7, cin=32, cout=40)
    _case(eng, "3x3 s2 c64->64 16x24", n=2, h=16, w=24, cin=64, cout=64, stride=2)
    _case(eng, "1x1 gate c64", n=2, h=8, w=8, cin=64, cout=64, ksize=1, mode=_lib.EPI_GATE, want=("f32", "silu"))
    _case(eng, "1x1 gdn c64", n=2, h=8, w=8, cin=64, cout=64, ksize=1, mode=_lib.EPI_GDN, want=("raw",))
    _case(eng, "3x3 shuffle c64->256", n=2, h=8, w=8, cin=64, cout=256, store=_lib.STORE_SHUFFLE_NHWC, want=("f32", "sq"))
    _case(eng, "3x3 final c64->12 nchw", n=2, h=8, w=8, cin=64, cout=12, store=_lib.STORE_SHUFFLE_NCHW)
    _case(eng, "3x3 res1(-1)+res2 passes1", n=2, h=8, w=8, cin=64, cout=64, passes=1, use_res1=True, res1_scale=-1.0,
          use_res2=True, want=("f32", "raw", "silu"))


def g_tc_basic():
    from mcquic_b200.engine import Engine
    eng = Engine("tcgen05")
    _case(eng, "tc 3x3 s1 c128 16x16 n1 p1", n=1, h=16, w=16, cin=128, cout=128, passes=1)
    _case(eng, "tc 3x3 s1 c128 16x16 n1 p3", n=1, h=16, w=16, cin=128, cout=128, passes=3)
    _case(eng, "tc 1x1 c128 16x16 n1 p3", n=1, h=16, w=16, cin=128, cout=128, ksize=1, passes=3)
    _case(eng, "tc 3x3 s1 c128 64x64 n4 p3 (multi-tile)", n=4, h=64, w=64, cin=128, cout=128, passes=3)
    _case(eng, "tc 3x3 s1 c128 64x64 n4 p1 (multi-tile)", n=4, h=64, w=64, cin=128, cout=128, passes=1)


def g_tc_shapes():
    from mcquic_b200 import _lib
    from mcquic_b200.engine import Engine
    eng = Engine("tcgen05")
    _case(eng, "tc 8x8 n4 (tn=2)", n=4, h=8, w=8, cin=128, cout=128)
    _case(eng, "tc 4x4 n5 (tn=8, ragged n)", n=5, h=4, w=4, cin=128, cout=128)
    _case(eng, "tc 2x2 n3", n=3, h=2, w=2, cin=128, cout=128)
    _case(eng, "tc 24x40 n2 (ragged h,w)", n=2, h=24, w=40, cin=128, cout=128)
    _case(eng, "tc 6x12 n3 (non-pow2)", n=3, h=6, w=12, cin=128, cout=128)
    _case(eng, "tc c192 16x16", n=2, h=16, w=16, cin=192, cout=192)
    _case(eng, "tc c192 16x16 p1", n=2, h=16, w=16, cin=192, cout=192, passes=1)
    _case(eng, "tc c64->64", n=2, h=16, w=16, cin=64, cout=64)
    _case(eng, "tc shuffle c128->512", n=2, h=16, w=16, cin=128, cout=512, store=_lib.STORE_SHUFFLE_NHWC, want=("f32", "sq"))
    _case(eng, "tc final c128->12 nchw p1", n=2, h=32, w=32, cin=128, cout=12, store=_lib.STORE_SHUFFLE_NCHW, passes=1)
    _case(eng, "tc final c128->12 nchw p3", n=2, h=32, w=32, cin=128, cout=12, store=_lib.STORE_SHUFFLE_NCHW, passes=3)


def g_tc_stride2():
    from mcquic_b200.engine import Engine
    eng = Engine("tcgen05")
    _case(eng, "tc 3x3 s2 c128 32x32->16x16 p3", n=2, h=32, w=32, cin=128, cout=128, stride=2)
    _case(eng, "tc 3x3 s2 c128 32x32->16x16 p1", n=2, h=32, w=32, cin=128, cout=128, stride=2, passes=1)
    _case(eng, "tc 3x3 s2 c128 8x8->4x4", n=3, h=8, w=8, cin=128, cout=128, stride=2)
    _case(eng, "tc 3x3 s2 c128 48x80", n=1, h=48, w=80, cin=128, cout=128, stride=2)


def g_tc_epi():
    from mcquic_b200 import _lib
    from mcquic_b200.engine import Engine
    eng = Engine("tcgen05")
    _case(eng, "tc gate 1x1", n=2, h=16, w=16, cin=128, cout=128, ksize=1, mode=_lib.EPI_GATE, want=("f32", "silu"))
    _case(eng, "tc gdn 1x1", n=2, h=16, w=16, cin=128, cout=128, ksize=1, mode=_lib.EPI_GDN, want=("raw",))
    _case(eng, "tc igdn 1x1", n=2, h=16, w=16, cin=128, cout=128, ksize=1, mode=_lib.EPI_IGDN, want=("raw",))
    _case(eng, "tc res1(-1)+res2 f32/raw/silu", n=2, h=16, w=16, cin=128, cout=128, use_res1=True, res1_scale=-1.0,
          use_res2=True, want=("f32", "raw", "silu")[:3] if False else ("f32", "raw"))
    _case(eng, "tc silu+sq planes p1", n=2, h=16, w=16, cin=128, cout=128, passes=1, want=("silu", "sq"))


def g_vq():
    import torch
    from mcquic_b200.engine import Engine
    eng = Engine()
    for (n, h, w, m, k, d) in [(2, 16, 16, 1, 8192, 128), (3, 7, 5, 6, 2048, 32), (1, 4, 4, 2, 100, 8), (2, 8, 8, 1, 512, 128)]:
        g = torch.Generator().manual_seed(1)
        x = (torch.randn(n, h, w, m * d, generator=g) * 0.15).cuda()
        cb = (torch.randn(m, k, d, generator=g) * 0.11).cuda()
        c2 = (cb ** 2).sum(-1).contiguous()
        hist = torch.zeros(m * k, dtype=torch.int32, device="cuda")
        codes, lg = eng.vq_assign(x, cb, c2, n, h, w, logits=True, hist=hist)
        xd = x.double().reshape(n * h * w, m, d)
        dist = ((xd ** 2).sum(-1)[..., None] + (cb.double() ** 2).sum(-1)[None]) - 2 * torch.einsum("pmd,mkd->pmk", xd, cb.double())
        ref = dist.argmin(-1).reshape(n, h * w, m).permute(0, 2, 1).reshape(n, m, h, w)
        mism = int((ref != codes).sum())
        lref = (-dist / k ** 0.5).reshape(n, h, w, m, k).permute(0, 3, 1, 2, 4)
        lerr = float((lg.double() - lref).abs().max())
        hok = bool((hist.reshape(m, k).sum(-1) == n * h * w).all())
        deq = eng.vq_dequant(codes, cb, {"f32", "silu"})
        dref = torch.gather(cb[None].expand(n * h * w, m, k, d), 2, codes.permute(0, 2, 3, 1).reshape(n * h * w, m, 1, 1).expand(-1, -1, 1, d)).reshape(n, h, w, m * d)
        derr = float((deq.f32 - dref).abs().max())
        print(f"  vq n{n} {h}x{w} m{m} k{k} d{d}: mismatches {mism}/{codes.numel()} logit err {lerr:.3e} hist ok {hok} dequant err {derr}", flush=True)


def g_accuracy():
    """error of the 3-pass tcgen05 conv vs fp64 on fp32 inputs, next to cuDNN fp32 and the SIMT kernel"""
    import torch
    import torch.nn.functional as F
    from convcase import make_planes
    from mcquic_b200.engine import Act, Engine, pack_conv
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    n, h, w, c = 4, 32, 32, 128
    x = torch.randn(n, h, w, c, generator=g).cuda()
    wt = ((torch.rand(c, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
    b = torch.zeros(c).cuda()
    exact = F.conv2d(x.double().permute(0, 3, 1, 2), wt.double(), None, padding=1).permute(0, 2, 3, 1)
    cudnn = F.conv2d(x.permute(0, 3, 1, 2), wt, None, padding=1).permute(0, 2, 3, 1)
    scale = float(exact.abs().max())
    print(f"  cuDNN fp32          max err/scale {float((cudnn.double()-exact).abs().max())/scale:.3e} rms {float((cudnn.double()-exact).pow(2).mean().sqrt())/scale:.3e}")
    for impl in ("simt", "tcgen05"):
        for passes in (3, 1):
            eng = Engine(impl)
            eng.passes = passes
            pc = pack_conv(wt, b, 1, 0, "cuda")
            out = eng.conv(pc, make_planes(x, passes), Act(n, h, w, c), {"f32"})
            d = out.f32.double() - exact
            print(f"  {impl:8s} passes={passes}  max err/scale {float(d.abs().max())/scale:.3e} rms {float(d.pow(2).mean().sqrt())/scale:.3e} mean(signed) {float(d.mean())/scale:.3e}", flush=True)


def _model(impl, C=128, M=1, K=(8192, 2048, 512), N=2, H=256, W=256):
    import torch
    from mcquic_b200 import Compressor
    from mcquic_b200.utils.synthetic import synthetic_state_dict
    from oracle import mcquic_oracle as O
    sd = synthetic_state_dict(C, M, list(K), seed=0)
    torch.manual_seed(0)
    x = torch.rand(N, 3, H, W) * 2 - 1
    t = time.time()
    oc, marg = O.encode(sd, x, with_margin=True)
    ox = O.decode(sd, oc)
    print(f"  oracle (cpu) encode+decode {time.time()-t:.2f}s", flush=True)
    model = Compressor(C, M, list(K)).eval()
    model.load_state_dict(sd)
    model = model.cuda()
    model.set_impl(impl)
    codes = model.encode(x.cuda())
    torch.cuda.synchronize()
    for lv, (a, b) in enumerate(zip(codes, oc)):
        mism = (a.cpu() != b)
        print(f"  [{impl}] level {lv} {tuple(a.shape)} mismatches {int(mism.sum())}/{a.numel()} min margin {float(marg[lv].min()):.2e} margins@mismatch {marg[lv][mism].tolist()[:6]}", flush=True)
    for passes in (1, 3):
        model.decode_passes = passes
        xh = model.decode([c.cuda() for c in oc]).cpu()
        print(f"  [{impl}] decode passes={passes} max abs err {float((xh-ox).abs().max()):.3e} (pixel range {float(ox.min()):.3f}..{float(ox.max()):.3f})", flush=True)
    print("  device flag", model.engine.lib.mcq_device_error_flag())


def g_model_simt():
    _model("simt", N=1)


def g_model_tc():
    _model("tcgen05", N=2)
    _model("tcgen05", C=192, M=6, K=(2048, 2048, 2048), N=1, H=128, W=256)


def g_timing():
    import torch
    from convcase import make_planes
    from mcquic_b200 import Compressor, _lib
    from mcquic_b200.engine import Act, Engine, pack_conv
    eng = Engine("tcgen05")
    g = torch.Generator().manual_seed(0)

    def tconv(n, h, w, cin, cout, passes, stride=1, ksize=3, store=0, reps=10):
        x = torch.randn(n, h, w, cin, generator=g).cuda()
        wt = ((torch.rand(cout, cin, ksize, ksize, generator=g) * 2 - 1) / (ksize * ksize * cin) ** 0.5).cuda()
        pc = pack_conv(wt, torch.zeros(cout).cuda(), stride, store, "cuda")
        eng.passes = passes
        a = make_planes(x, passes)
        act = Act(n, h, w, cin)
        for _ in range(3):
            eng.conv(pc, a, act, {"f32", "silu"} if store == 0 else {"f32"})
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            eng.conv(pc, a, act, {"f32", "silu"} if store == 0 else {"f32"})
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        fl = 2.0 * n * (h // stride) * (w // stride) * cout * cin * ksize * ksize
        print(f"  conv n{n} {h}x{w} c{cin}->{cout} k{ksize} s{stride} passes={passes}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s(alg)", flush=True)

    for passes in (3, 1):
        tconv(64, 128, 128, 128, 128, passes)
        tconv(64, 64, 64, 128, 128, passes)
        tconv(64, 32, 32, 128, 128, passes)
        tconv(64, 16, 16, 128, 128, passes)
        tconv(64, 8, 8, 128, 128, passes)
        tconv(64, 4, 4, 128, 128, passes)
        tconv(64, 128, 128, 128, 128, passes, stride=2)
        tconv(64, 64, 64, 128, 512, passes, store=_lib.STORE_SHUFFLE_NHWC)
        tconv(64, 64, 64, 128, 128, passes, ksize=1)

    from mcquic_b200.utils.synthetic import synthetic_state_dict
    model = Compressor(128, 1, [8192, 2048, 512]).eval()
    model.load_state_dict(synthetic_state_dict(128, 1, [8192, 2048, 512], seed=0))
    model = model.cuda()
    for N in (8, 64):
        x = (torch.rand(N, 3, 256, 256, generator=g) * 2 - 1).cuda()
        for _ in range(2):
            codes = model.encode(x)
            xh = model.decode(codes)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(3):
            codes = model.encode(x)
        e[1].record()
        for _ in range(3):
            xh = model.decode(codes)
        e[2].record()
        torch.cuda.synchronize()
        te, td = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
        print(f"  model N={N}: encode {te:.2f} ms decode {td:.2f} ms -> {N*256*256/(te+td)/1e3:.1f} MPix/s (eager launches, no graph)", flush=True)


def g_halo():
    """halo kernel variants; configuration = library options given as MCQ_OPTIONS="halo=1,halo_cl=2,..." (applied by this
    script through mcq_set_option; the library itself never reads the environment)"""
    import torch
    from convcase import make_planes
    from mcquic_b200 import _lib
    from mcquic_b200.engine import Act, Engine, pack_conv
    print("  options", os.environ.get("MCQ_OPTIONS", ""), flush=True)
    _lib.apply_options(os.environ.get("MCQ_OPTIONS", ""))
    eng = Engine("tcgen05")
    for passes in (1, 3):
        _case(eng, f"halo 16x16 n1 p{passes}", n=1, h=16, w=16, cin=128, cout=128, passes=passes)
        _case(eng, f"halo 64x64 n4 p{passes}", n=4, h=64, w=64, cin=128, cout=128, passes=passes)
        _case(eng, f"halo 24x40 n3 p{passes} ragged", n=3, h=24, w=40, cin=128, cout=128, passes=passes)
        _case(eng, f"halo 32x32 n2 c192 p{passes}", n=2, h=32, w=32, cin=192, cout=192, passes=passes)
        _case(eng, f"halo shuffle c128->512 p{passes}", n=2, h=16, w=16, cin=128, cout=512, passes=passes,
              store=_lib.STORE_SHUFFLE_NHWC, want=("f32", "sq"))
    g = torch.Generator().manual_seed(0)
    for passes in (3, 1):
        for (n, h, w, cin, cout) in [(64, 128, 128, 128, 128), (64, 64, 64, 128, 128), (64, 64, 64, 128, 512), (64, 32, 32, 128, 128), (64, 16, 16, 128, 128), (64, 8, 8, 128, 128), (64, 4, 4, 128, 128), (64, 128, 128, 128, 12)]:
            x = torch.randn(n, h, w, cin, generator=g).cuda()
            wt = ((torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) / (9 * cin) ** 0.5).cuda()
            store = _lib.STORE_SHUFFLE_NHWC if cout == 512 else (_lib.STORE_SHUFFLE_NCHW if cout == 12 else 0)
            pc = pack_conv(wt, torch.zeros(cout).cuda(), 1, store, "cuda")
            eng.passes = passes
            a = make_planes(x, passes)
            act = Act(n, h, w, cin)
            want = {"f32", "silu"} if store == 0 else {"f32"}
            res = torch.randn(n, h, w, cout, generator=g).cuda() if (store == 0 and os.environ.get("SELFTEST_RES", "1") == "1") else None
            gr = torch.cuda.CUDAGraph()
            for _ in range(2):
                eng.conv(pc, a, act, want, res1=res)
            torch.cuda.synchronize()
            with torch.cuda.graph(gr):
                for _ in range(10):
                    o = eng.conv(pc, a, act, want, res1=res)
            gr.replay()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            gr.replay()
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            fl = 2.0 * n * h * w * cout * cin * 9
            print(f"  conv n{n} {h}x{w} c{cin}->{cout} passes={passes}: {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s(alg)", flush=True)
    print("  flag", eng.lib.mcq_device_error_flag())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "halo_sweep":
        combos = [("0", "10", "0", "1", "0"), ("1", "10", "0", "1", "0"), ("1", "10", "0", "2", "0"), ("1", "10", "0", "2", "1"), ("0", "10", "0", "1", "1")]
        for (on, pitch, base, cl, skip) in combos:
            env = dict(os.environ, MCQ_OPTIONS=f"halo={on},halo_pitch={pitch},halo_base={base},halo_cl={cl},epi_skip={skip}")
            print(f"==== halo={on} pitch={pitch} base={base} cl={cl} epi_skip={skip}", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "halo"], timeout=300, capture_output=True, text=True, env=env)
                print(r.stdout[-5000:], end="")
                if r.returncode != 0:
                    print(f"!! exit {r.returncode}\n{r.stderr[-1500:]}")
            except subprocess.TimeoutExpired as e:
                print("!! TIMEOUT", (e.stdout or b"")[-2000:])
        return
    if len(sys.argv) > 1:
        import faulthandler
        faulthandler.dump_traceback_later(150, exit=True)
        for name in sys.argv[1:]:
            print(f"== {name}", flush=True)
            globals()["g_" + name]()
        return
    for name in GROUPS:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=600, capture_output=True, text=True)
            print(r.stdout[-6000:], end="")
            if r.returncode != 0:
                print(f"!! group {name} exit {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"!! group {name} TIMEOUT\n{(e.stdout or b'')[-3000:]}")
        print(f"-- {name} took {time.time()-t:.1f}s", flush=True)


if __name__ == "__main__":
    main()
