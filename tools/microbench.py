"""Quick device-timed microbenchmarks (config 5 of BASELINE.json): NTT 2^22, MSM 2^20, plus
the k=17 prover shapes.  Development tool; bench.py is the contract."""
import sys, os, time, argparse
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("halo2-rsa_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import b2rsa
import bn254 as O
from util import fr_to_np, random_fr_np, g1_to_np


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)


def random_bases(n, seed=1):
    """n pseudo-random distinct points: (a_i)G is slow in python; use multiples via C-free trick:
    take P_i = (i+1)G (running add, batch normalise) - fine for timing and closed-form checks."""
    import cpu_oracle as CO
    return CO.g1_multiples(n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt", type=int, nargs="*", default=[17, 19, 22])
    ap.add_argument("--msm", type=int, nargs="*", default=[17, 20])
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    ctx = b2rsa.Context(0)
    # torch's default stream has handle 0 (== "own stream" for b2r_ctx_set_stream): time on a side stream
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    ctx.set_stream(side.cuda_stream)
    for L in a.ntt:
        n = 1 << L
        x = torch.from_numpy(random_fr_np(n, L).view(np.int64)).cuda()
        w = fr_to_np([O.omega_for(L)])[0]
        best, avg = timeit(lambda: ctx.ntt_batch_dev(x.data_ptr(), 1, w, L))
        gb = 2 * n * 32 / 1e9
        print(f"ntt 2^{L}: best {best:.3f} ms avg {avg:.3f} ms  -> {gb/best*1e3:.1f} GB/s algorithmic", flush=True)
        if L <= 19:
            xb = torch.from_numpy(random_fr_np(n * a.batch, L + 100).view(np.int64)).cuda()
            best, avg = timeit(lambda: ctx.ntt_batch_dev(xb.data_ptr(), a.batch, w, L))
            print(f"ntt 2^{L} x{a.batch}: best {best:.3f} ms ({best/a.batch:.3f} ms each) -> {gb*a.batch/best*1e3:.1f} GB/s", flush=True)
    for L in a.msm:
        n = 1 << L
        t0 = time.time(); pts = random_bases(n); t1 = time.time()
        bs = ctx.bases_register(pts); t2 = time.time()
        print(f"msm 2^{L}: bases gen {t1-t0:.1f}s register {t2-t1:.2f}s", flush=True)
        sc = torch.from_numpy(random_fr_np(n, 0x5EED).view(np.int64)).cuda()
        out = torch.zeros(8 * a.batch, dtype=torch.int64, device="cuda")
        best, avg = timeit(lambda: ctx.msm_batch_dev(bs, sc.data_ptr(), 1, n, out.data_ptr(), uniform=True))
        gb = n * 96 / 1e9
        print(f"msm 2^{L} uniform: best {best:.3f} ms avg {avg:.3f} -> {gb/best*1e3:.1f} GB/s algorithmic", flush=True)
        ctx.profile_enable(True); ctx.profile_dump(clear=True)
        ctx.msm_batch_dev(bs, sc.data_ptr(), 1, n, out.data_ptr(), uniform=True)
        print("   per kernel (ms):", {k: round(v[0], 3) for k, v in ctx.profile_dump(clear=True).items()}, flush=True)
        ctx.profile_enable(False)
        best, avg = timeit(lambda: ctx.msm_batch_dev(bs, sc.data_ptr(), 1, n, out.data_ptr(), uniform=False))
        print(f"msm 2^{L} (no hint): best {best:.3f} ms avg {avg:.3f}", flush=True)
        if L <= 17:
            scb = torch.from_numpy(random_fr_np(n * a.batch, 7).view(np.int64)).cuda()
            best, avg = timeit(lambda: ctx.msm_batch_dev(bs, scb.data_ptr(), a.batch, n, out.data_ptr(), uniform=True))
            print(f"msm 2^{L} x{a.batch} uniform: best {best:.3f} ms ({best/a.batch:.3f} each)", flush=True)
            ctx.profile_enable(True); ctx.profile_dump(clear=True)
            ctx.msm_batch_dev(bs, scb.data_ptr(), a.batch, n, out.data_ptr(), uniform=True)
            print("   per kernel (ms):", {k: round(v[0], 3) for k, v in ctx.profile_dump(clear=True).items()}, flush=True)
            ctx.profile_enable(False)
            # advice-like skew: 40% zero, 30% one, 20% bytes, 9% 64-bit, 1% full
            rng = np.random.default_rng(3)
            canon = np.zeros((n * a.batch, 4), dtype=np.uint64)
            sel = rng.random(n * a.batch)
            canon[(sel >= 0.4) & (sel < 0.7), 0] = 1
            m = (sel >= 0.7) & (sel < 0.9); canon[m, 0] = rng.integers(0, 256, size=int(m.sum()), dtype=np.uint64)
            m = (sel >= 0.9) & (sel < 0.99); canon[m, 0] = rng.integers(0, 1 << 63, size=int(m.sum()), dtype=np.uint64)
            full = random_fr_np(n * a.batch, 9); m = sel >= 0.99; canon[m] = full[m]
            # canonical -> montgomery on device is not exposed; emulate: scalars here are "already montgomery" values,
            # the kernel converts from_mont, so skew must be in canonical domain: build montgomery of small ints in python (slow) for one vector
            small = [int(v) for v in canon[:n, 0]]
            mont = fr_to_np(small)
            mont[m[:n]] = full[:n][m[:n]]
            sk = torch.from_numpy(np.tile(mont, (a.batch, 1)).view(np.int64)).cuda()
            best, avg = timeit(lambda: ctx.msm_batch_dev(bs, sk.data_ptr(), a.batch, n, out.data_ptr()))
            print(f"msm 2^{L} x{a.batch} advice-like skew: best {best:.3f} ms ({best/a.batch:.3f} each)", flush=True)
            ctx.profile_enable(True); ctx.profile_dump(clear=True)
            ctx.msm_batch_dev(bs, sk.data_ptr(), a.batch, n, out.data_ptr())
            print("   per kernel (ms):", {k: round(v[0], 3) for k, v in ctx.profile_dump(clear=True).items()}, flush=True)
            ctx.profile_enable(False)
        bs.free()
    print("launches", ctx.launch_count)

if __name__ == "__main__":
    main()
