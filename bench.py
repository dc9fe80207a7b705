"""bench.py - RSA-2048 pkcs1v15 proofs/sec at k=17 on N B200s (BASELINE.json configs[1]).

One "step" = ONE COMPLETE PROOF for each of 64 synthetic RSA-2048 instances per GPU, i.e. the reference bench's
create_proof (benches/bench.rs:319-331) for a batch: witness synthesis -> 5 advice commitments -> lookup
permutations (10 commitments) -> permutation / lookup grand products (7) + random polynomial -> 22 x
(lagrange_to_coeff 2^17, coeff_to_extended 2^19) -> quotient on the 2^19 coset -> 4 h commitments ->
58 evaluations -> GWC multiopen (4 commitments): 31 MSMs of 2^17 points and 45 NTTs per proof, with the
Blake2b transcripts on the host between the phases.  Output = 64 proofs of 2848 bytes.

  value : proofs / s with the instance inputs resident in HBM (b2r_rsa_prove_batch_dev; CUDA events)
  e2e   : the same through the reference-facing C-ABI call b2r_rsa_prove_batch with pinned HOST inputs
          (h2d inside the timed region; the proofs always come back to the host)
  --impl reference : the CPU arm.  The reference is Rust and no cargo/rustc exists in this image, so this
          times oracle/ - this repo's C restatement of the reference's CPU prover - on all host cores for ONE
          COMPLETE PROOF per step: sequential Circuit::synthesize (oracle/rsa_witness.c) followed by the whole
          create_proof (oracle/plonk_prover.c: 31 best_multiexp, 45 best_fft, lookups, grand products, quotient,
          evaluations, GWC multiopen, Blake2b transcript), the same work the GPU arm does per instance.
  extra : BASELINE configs[4] (standalone MSM 2^20 / NTT 2^22) measured after the timed region with CUDA events.

Multi-GPU (torchrun, one rank per GPU): instances are independent, each rank proves its own 64, then ONE
NCCL all_gather of the device-resident commitment block (31 affine points per proof, 127 KB per rank); scaling = weak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in ("halo2-rsa_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

import numpy as np  # noqa: E402

BITS, K, EXT_K, BATCH = 2048, 17, 19, 64
NCOL = 5
MSM_PER_PROOF, INTT_PER_PROOF, COSET_PER_PROOF = 31, 22, 22
METRIC = "RSA-2048 pkcs1v15 proofs/sec at k=17"
WORKLOAD = ("rsa2048_e65537_k17_batch64_per_gpu: full create_proof per instance (witness, 31 MSM 2^17, 22 iNTT 2^17, "
            "22 coset NTT 2^19 + 1 coset iNTT 2^19, lookups, permutation, quotient, 58 evals, GWC multiopen, Blake2b transcript)")
MSM_BYTES_PER_TERM = 96  # 32 B scalar + 64 B affine base
FULL_MSM_PER_PROOF, MSM_WINDOWS = 16, 16  # grand products 7, random poly 1, h pieces 4, GWC witnesses 4; ceil(255 / 16) windows
CSRC = os.path.join(ROOT, "halo2-rsa_b200", "csrc")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")


def kernel_revision(files):
    """sha1 over the kernel sources a DRAM-traffic capture depends on (tools/traffic_from_ncu.py stamps the same value)"""
    import hashlib
    h = hashlib.sha1()
    for f in files:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()[:16]


def recorded_traffic(name, files):
    """roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch from the last `ncu --set full` capture
    (profiles/traffic.json, written by tools/traffic_from_ncu.py).  The record carries the revision of the kernel
    sources it was captured at; if they changed since, the number is refused (null + note) instead of going stale."""
    try:
        rec = json.load(open(TRAFFIC_FILE))[name]
    except Exception:
        return None, {"traffic_note": "no ncu capture recorded for this kernel (profiles/traffic.json)"}
    now = kernel_revision(files)
    if rec.get("kernel_revision") != now:
        return None, {"traffic_note": f"stale capture refused: recorded at kernel revision {rec.get('kernel_revision')}, sources are now {now}"}
    return rec["dram_bytes_per_launch"], {"traffic_algorithmic_bytes_of_that_launch": rec.get("algorithmic_bytes_per_launch"),
                                          "traffic_source": rec.get("source"), "traffic_kernel_revision": now}


MSM_SOURCES = ("msm.cu", "ec.cuh", "field.cuh")
NTT_SOURCES = ("ntt.cu", "field.cuh")


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_step(samples, threads):
    """the CPU arm (oracle 'port'): seconds for `samples` COMPLETE proofs, one after the other, the way the reference's
    bench proves (benches/bench.rs:319-331): single-threaded synthesize (as the reference's SimpleFloorPlanner pass),
    then create_proof on `threads` host threads (oracle/plonk_prover.c)."""
    import cpu_oracle as CO
    import plonk as PL
    import rsa_fixtures as RF
    st = cpu_port_state(threads)
    t0 = time.perf_counter()
    for i in range(samples):
        n, s, h = RF.instance(BITS, i)
        t = CO.RsaTable(BITS, K)
        ok = t.synthesize(RF.limbs64(n, BITS // 64), RF.limbs64(s, BITS // 64), RF.limbs64(h, 4))
        assert ok == 1
        adv = t.advice()
        t.free()
        proof, _ = PL.create_proof_c(st["ka"], st["srs"], adv, 0xB200, proof_index=i, nonce=st["nonce"], threads=threads)
        st["nonce"] += 1
        st["last_proof"] = proof
    return time.perf_counter() - t0


_cpu_state = None


def cpu_port_state(threads):
    """SRS + keygen of the CPU arm (outside every timed region, like ParamsKZG::setup / keygen_pk in the reference's bench)"""
    global _cpu_state
    if _cpu_state is None:
        import cpu_oracle as CO
        import plonk as PL
        CO.build()
        srs = PL.Srs(K)
        ka = PL.keygen_arrays(PL.circuit_layout(BITS, K), srs, threads)
        _cpu_state = {"srs": srs, "ka": ka, "nonce": 0, "last_proof": None}
    return _cpu_state


def run_reference(args):
    """--impl reference: CPU arm on rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cpu_port_state(threads)
    sample = 1
    st = cpu_port_state(threads)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_port_step(sample, threads)
    import plonk as PL
    vk = PL.vk_from_commitments(K, st["ka"]["fixed_commitments"], st["ka"]["sigma_commitments"], st["ka"]["transcript_repr"])
    assert PL.verify_proof(vk, st["srs"].s, st["last_proof"]), "CPU arm produced a proof its verifier rejects"
    t = 0.0
    for _ in range(args.steps):
        t += cpu_port_step(sample, threads)
    val = sample * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": 1, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (BN254 Fr/Fq Montgomery, integer)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": f"{sample} proof (bounded sample of the 64-instance batch)"},
        "cpu_baseline": {"value": val, "unit": "proofs/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} complete proof per step x {args.steps} steps: synthesize on 1 thread, then create_proof (31 MSM, 45 FFT, lookups, grand products, quotient, evaluations, multiopen, transcript) on {threads} threads"},
        "e2e": {"value": val, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust (no cargo/rustc in this image): timed arm is oracle/ - the C restatement of its CPU prover (rsa_witness.c + plonk_prover.c)",
    }
    print(json.dumps(line), flush=True)


def config5_microbench(ctx, torch, stream):
    """BASELINE configs[4]: standalone BN254 MSM 2^20 (bases (i+1) G, uniform seeded scalars) and forward NTT 2^22
    (uniform input, omega = ROOT_OF_UNITY^(2^6)), device resident, CUDA events on the library's stream, best and mean of
    5 after 2 warm-ups; outside the timed region of the headline metric.  GB/s = algorithmic bytes / time."""
    import cpu_oracle as CO
    import bn254 as O
    from util import fr_to_np, random_fr_np
    peak = (measured_peaks() or {}).get("hbm_gbs", 6650.0)

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts), sum(ts) / len(ts)

    out = {}
    n = 1 << 20
    bs = ctx.bases_register(CO.g1_multiples(n))
    sc = torch.from_numpy(random_fr_np(n, 0x5EED).view(np.int64)).cuda()
    res = torch.zeros(8, dtype=torch.int64, device="cuda")
    best, mean = timeit(lambda: ctx.msm_batch_dev(bs, sc.data_ptr(), 1, n, res.data_ptr(), uniform=True))
    gb = n * MSM_BYTES_PER_TERM / 1e9
    out["msm_2p20"] = {"ms_best": best, "ms_mean": mean, "algorithmic_GBps": gb / best * 1e3, "frac_of_hbm_peak": gb / best * 1e3 / peak,
                       "terms_per_s": n / best * 1e3, "workload": "2^20 uniform scalars, bases (i+1)G resident (c = 16 window table)"}
    bs.free()
    del sc
    log_n = 22
    m = 1 << log_n
    a = torch.from_numpy(random_fr_np(m, 0x5EED + 1).view(np.int64)).cuda()
    w = fr_to_np([O.omega_for(log_n)])[0]
    best, mean = timeit(lambda: ctx.ntt_batch_dev(a.data_ptr(), 1, w, log_n))
    gb = 2 * m * 32 / 1e9
    out["ntt_2p22"] = {"ms_best": best, "ms_mean": mean, "algorithmic_GBps": gb / best * 1e3, "frac_of_hbm_peak": gb / best * 1e3 / peak,
                       "workload": "forward NTT 2^22 in place, natural order in and out"}
    out["peak_GBps"] = peak
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--workload", default="rsa2048", choices=["rsa2048", "rsa4096"],
                    help="rsa2048 = BASELINE configs[1] (the headline line, default); rsa4096 = configs[2] (RSA-4096, k = 18, batch 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the BASELINE configs[4] microbench (MSM 2^20 / NTT 2^22)")
    args = ap.parse_args()
    global BITS, K, EXT_K, BATCH, WORKLOAD
    if args.workload == "rsa4096":
        BITS, K, EXT_K, BATCH = 4096, 18, 20, 32
        WORKLOAD = WORKLOAD.replace("rsa2048_e65537_k17_batch64_per_gpu", "rsa4096_e65537_k18_batch32_per_gpu").replace("2^17", "2^18").replace("2^19", "2^20")
    if args.batch is None:
        args.batch = BATCH
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import b2rsa
    import rsa_fixtures as RF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # ONE JSON line on stdout: libraries write banners to file descriptor 1 (NCCL prints "NCCL version ..." there whenever
    # NCCL_DEBUG is VERSION / WARN / INFO in the environment), so fd 1 is pointed at stderr for the whole run and the
    # result line goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)  # torch's default stream has handle 0; the library needs a real stream to share

    import b2rsa.shard as shard
    ctx = b2rsa.Context(local)
    ctx.set_stream(side.cuda_stream)
    batch, n = args.batch, 1 << K
    prog = ctx.rsa_program(BITS, K)
    import bn254 as O
    from util import fr_to_np
    g, gl = ctx.srs_setup(K, fr_to_np([O.srs_secret(K)])[0])
    pk = ctx.rsa_keygen(prog, g, gl)
    pb = pk.proof_bytes

    # rank r proves instances [r*batch, (r+1)*batch)
    mine = shard.instance_range(rank, world, world * batch)
    nl, sl, hl = RF.batch(BITS, batch, start=mine.start)
    h_n = torch.from_numpy(nl.view(np.int64)).pin_memory()
    h_s = torch.from_numpy(sl.view(np.int64)).pin_memory()
    h_h = torch.from_numpy(hl.view(np.int64)).pin_memory()
    d_n, d_s, d_h = h_n.to(dev), h_s.to(dev), h_h.to(dev)
    h_proofs = torch.zeros(batch * pb, dtype=torch.uint8).pin_memory()
    h_status = torch.zeros(batch, dtype=torch.uint8).pin_memory()
    ncm = ctx.last_commitments_info()[1]                                          # 31 commitments per proof
    d_cm = torch.zeros((batch, ncm * 8), dtype=torch.int64, device=dev)           # this rank's device-resident commitment block
    seed = 0xB200 + rank

    gathered = {}

    def finish():
        if world > 1:
            # the commitments never leave the device: the library files them per phase (b2r_last_commitments), one
            # device-to-device copy on the shared stream, one NCCL all-gather
            ctx.last_commitments_dev(d_cm.data_ptr(), batch * ncm)
            gathered["all"] = shard.gather_commitments(d_cm, world * batch)

    def step_dev():
        pk.prove_batch_raw(d_n.data_ptr(), d_s.data_ptr(), d_h.data_ptr(), batch, seed, h_proofs.data_ptr(), h_status.data_ptr(),
                           inputs_on_device=True)
        finish()

    def step_e2e():
        pk.prove_batch_raw(h_n.data_ptr(), h_s.data_ptr(), h_h.data_ptr(), batch, seed, h_proofs.data_ptr(), h_status.data_ptr())
        finish()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(steps):
            fn()
        e1.record(side)
        barrier()
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    assert h_status.tolist() == [1] * batch, "synthetic signatures must verify"
    if rank == 0:  # the timed path's output is a real proof: check one against the oracle verifier
        import plonk as PL
        from util import np_to_fr, np_to_g1
        f, s_, t = pk.export_vk()
        vk = PL.vk_from_commitments(K, np_to_g1(f), np_to_g1(s_), np_to_fr(t.reshape(1, 4))[0])
        assert PL.verify_proof(vk, O.srs_secret(K), bytes(h_proofs[:pb].numpy())), "proof rejected by the oracle verifier"
        if world > 1:   # the gathered block holds every rank's commitments; this rank's first row = the points of its first proof
            full = gathered["all"]
            assert full.shape == (world * batch, ncm * 8) and bool((full[batch:].abs().sum(dim=1) != 0).all())
            pts = np_to_g1(full[0].cpu().numpy().view(np.uint64).reshape(-1, 8))
            raw = bytes(h_proofs[:pb].numpy())
            want = [raw[32 * i:32 * i + 32] for i in range(ncm - 4)] + [raw[pb - 128 + 32 * i:pb - 96 + 32 * i] for i in range(4)]
            assert [PL.compress_point(P) for P in pts] == want, "gathered commitments differ from the proof's group elements"

    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile_enable(True)
    ctx.profile_dump(clear=True)
    launches0 = ctx.launch_count
    ms = timed(step_dev, args.steps)
    launches = ctx.launch_count - launches0
    _, _, accum_entries = ctx.profile_read("msm_accum_entries")   # bucket entries = mixed additions actually performed
    prof = ctx.profile_dump(clear=True)
    ctx.profile_enable(False)
    clocks = sampler.stop()

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    assert h_status.tolist() == [1] * batch

    value = world * batch * args.steps / (ms / 1e3)
    e2e_value = world * batch * args.steps / (ms_e2e / 1e3)

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = config5_microbench(ctx, torch, side)

    if rank == 0:
        peaks = measured_peaks()
        peak = (peaks or {}).get("hbm_gbs", 6650.0)
        roof = None
        if "msm_accum_entries" in prof:
            kms, kcnt = prof["msm_accum_entries"]
            terms_per_launch = batch * MSM_PER_PROOF * n * args.steps / kcnt
            achieved = terms_per_launch * MSM_BYTES_PER_TERM / (kms / kcnt / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_accum_entries (MSM bucket accumulation)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": terms_per_launch * MSM_BYTES_PER_TERM,
                    "avg_launch_ms": kms / kcnt, "share_of_step": kms / ms,
                    "note": "integer-multiplier bound (254-bit field products), not HBM bound: see int_pipe and DESIGN.md section 4"}
            # the roofline that binds this kernel (profiles/r01_pipebench.md): IMAD.WIDE issues at 32 lanes/clk/SM on B200, a
            # Montgomery product is 128 of them, the squaring 100, the fused a*b - c*d 192; a mixed addition is 6 products +
            # 2 squarings + 1 fused pair (csrc/ec.cuh xyzz_madd_ls) = 1160 -> 148 SMs * 32 / 1160 additions per clock
            sm_clk = 1.965e9
            ceiling = 148 * 32 / 1160.0 * sm_clk / 1e9
            roof["int_pipe"] = {"unit": "G mixed additions/s", "achieved": accum_entries / (kms / 1e3) / 1e9, "peak": ceiling,
                                "frac": accum_entries / (kms / 1e3) / 1e9 / ceiling,
                                "additions_per_step": accum_entries / args.steps,
                                "how": "bucket entries counted by the library (zero digits and, for the grand-product columns, rows where the "
                                       "column does not change are skipped) / kernel time; peak = 148 SMs x 32 IMAD.WIDE lanes/clk / "
                                       "(6 products x 128 + 2 squarings x 100 + 1 fused product pair x 192 = 1160 IMAD.WIDE per mixed addition) at 1965 MHz"}
            captured_cfg = args.workload == "rsa2048" and batch == 64   # what profiles/traffic.json was captured on
            roof["traffic"], extra_t = recorded_traffic("k_accum_entries", MSM_SOURCES) if captured_cfg else \
                (None, {"traffic_note": "profiles/traffic.json is a capture of the rsa2048 / batch 64 step"})
            roof.update(extra_t)
            if "ntt_pass" in prof:
                # second entry: the NTT passes, the kernel closest to being HBM relevant.  Algorithmic bytes per transform
                # = 2 * n * 32 B; `units` recorded by the library = elements per pass launch.
                nms, ncnt = prof["ntt_pass"]
                ntt_alg = args.steps * batch * (INTT_PER_PROOF * 2 * n * 32 + (COSET_PER_PROOF + 1) * 2 * (n << (EXT_K - K)) * 32)
                ntt_ach = ntt_alg / (nms / 1e3) / 1e9
                tr, extra_n = recorded_traffic("k_ntt_pass", NTT_SOURCES) if captured_cfg else \
                    (None, {"traffic_note": "profiles/traffic.json is a capture of the rsa2048 / batch 64 step"})
                roof["ntt"] = {"bound": "hbm", "kernel": "k_ntt_pass (all passes of the 45 transforms per proof)", "achieved": ntt_ach, "peak": peak, "unit": "GB/s",
                               "frac": ntt_ach / peak, "traffic": tr, "ms_per_step": nms / args.steps, "launches_per_step": ncnt / args.steps,
                               "share_of_step": nms / ms, "algorithmic_bytes_per_step": ntt_alg / args.steps}
                roof["ntt"].update(extra_n)
                if tr is not None:
                    # the capture covers whole steps' launches: per-step DRAM traffic = mean per launch x launches per step; it is
                    # the algorithmic 2 * n * 32 B per transform times the number of passes (2 at 2^17, 3 at 2^19): no re-reads
                    roof["ntt"]["traffic_per_step"] = tr * ncnt / args.steps
                    roof["ntt"]["traffic_over_algorithmic"] = tr * ncnt / ntt_alg
        line = {
            "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (BN254 Fr/Fq Montgomery, integer)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * batch, "proof_bytes": pb,
                       "parallelism": f"instances sharded x{world}, 1 all_gather of the device-resident commitments (31 points per proof)",
                       "l2": "inputs larger than L2 (the per-step polynomial arena is 17 GB)"},
            "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": int(h_n.numel() + h_s.numel() + h_h.numel()) * 8,
                    "d2h_bytes_per_step": int(h_proofs.numel()) + int(h_status.numel()), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        }
        if extra:
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_port_state(threads)
            sample = 2
            t = cpu_port_step(sample, threads)
            line["cpu_baseline"] = {"value": sample / t, "unit": "proofs/s", "cores": threads, "kind": "port",
                                    "sample": f"{sample} complete proofs of the same workload, one after the other: synthesize on 1 thread + the whole create_proof (oracle/plonk_prover.c) on {threads} threads ({t:.1f} s)"}
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
