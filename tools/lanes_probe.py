"""Development probe: do two proof groups in flight (two contexts = two streams + two scratch sets, one proving key)
beat one group of twice the size?  Upper bound for an in-library two-lane prover.  Not part of the product."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("halo2-rsa_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ctypes as C
import torch
import b2rsa
import bn254 as O
import rsa_fixtures as RF
from util import fr_to_np

BITS, K = 2048, 17


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = 4
    ctx = b2rsa.Context(0)
    prog = ctx.rsa_program(BITS, K)
    g, gl = ctx.srs_setup(K, fr_to_np([O.srs_secret(K)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    pb = pk.proof_bytes
    nl, sl, hl = RF.batch(BITS, total)

    def run(lanes):
        per = total // lanes
        ctxs = [ctx] + [b2rsa.Context(0) for _ in range(lanes - 1)]
        bufs = []
        for i in range(lanes):
            sl_ = slice(i * per, (i + 1) * per)
            hn = torch.from_numpy(nl[sl_].copy().view(np.int64)).pin_memory()
            hs = torch.from_numpy(sl[sl_].copy().view(np.int64)).pin_memory()
            hh = torch.from_numpy(hl[sl_].copy().view(np.int64)).pin_memory()
            pr = torch.zeros(per * pb, dtype=torch.uint8).pin_memory()
            st = torch.zeros(per, dtype=torch.uint8).pin_memory()
            bufs.append((hn, hs, hh, pr, st))

        def work(i, reps):
            c = ctxs[i]
            hn, hs, hh, pr, st = bufs[i]
            for _ in range(reps):
                rc = c.lib.b2r_rsa_prove_batch(c.h, pk.h, C.c_void_p(hn.data_ptr()), C.c_void_p(hs.data_ptr()), C.c_void_p(hh.data_ptr()), per,
                                               0xB200, C.c_void_p(pr.data_ptr()), C.c_void_p(st.data_ptr()))
                assert rc == 0, c.lib.b2r_last_error(c.h)
            assert bytes(st.numpy()) == b"\x01" * per

        def go(reps):
            th = [threading.Thread(target=work, args=(i, reps)) for i in range(lanes)]
            for t in th: t.start()
            for t in th: t.join()
        go(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        go(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"lanes {lanes}: {total * steps / dt:.2f} proofs/s ({dt / steps * 1e3:.1f} ms per {total} proofs)", flush=True)
        return bufs[0][3].numpy()[:pb].copy()

    a = run(1)
    b = run(2)
    run(1)
    run(2)
    if total >= 96:
        run(3)


if __name__ == "__main__":
    main()
