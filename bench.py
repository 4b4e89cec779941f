#!/usr/bin/env python
"""bench.py -- zerocaf hot path on B200 (and the reference-equivalent CPU path beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on rank 0.  A "step" is one pass of BASELINE config 2 over one batch: 2^24 FieldElement pairs ->
prod = a*b and sq = a^2, both fully reduced (2 x 2^24 253-bit field multiplications), one kernel launch.
  value   field-muls/s, whole job, inputs resident in HBM (weak scaling: every rank owns its own 2^24 batch, the path is
          element-wise and has no exchange step)
  e2e     the same metric through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside the timed region;
          the library pipelines the call in 16 MiB chunks: H2D of chunk k+1, kernel on k, D2H of k-1 overlap)
  roofline  algorithmic bytes (128 B per element pair: 2 x 32 in, 2 x 32 out) / kernel time vs measured HBM
  extra   config 3 (2^22 point add / double), config 4 (2^20 scalar-mul), config 5 (2^20-point MSM, window 16, sharded by
          bucket-window over the N ranks with one NCCL all-gather; points resident and prepared once, the per-call
          figure with the operand pass inside is ms_per_msm_unprepared) as secondary keys of the same line.
`--impl reference` times the CPU restatement of the reference's own algorithm (oracle/, kind "port": there is no Rust
toolchain in this image) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FIELD = 1 << 24      # config 2
N_POINT = 1 << 22      # config 3
N_SMUL = 1 << 20       # config 4
N_MSM = 1 << 20        # config 5
MSM_WINDOW = 16
BYTES_PER_PAIR = 128   # algorithmic bytes of one config-2 unit (canonical 32-byte encodings)
METRIC = "253-bit field-muls/s (config 2: 2^24 mul+square+reduce per step)"
UNIT = "field-muls/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing more than half of the highest power seen
        if pw:
            thr = 0.5 * max(pw)
            load = [s_ for s_, p_ in zip(sm, pw) if p_ >= thr] or sm
        else:
            load = sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference algorithm, all host threads, bounded sample
# ------------------------------------------------------------------------------------------------------------
_cpu_cache = {}


def cpu_field_rate(n_sample, threads, repeats=1):
    """Config 2 on the CPU port: 2 x n_sample field multiplications per call.  Inputs and the two output arrays are built
    and touched OUTSIDE the timed region (no page faults, no allocation inside it); the timed call is the arithmetic plus
    the pthread fork/join of one parallel_for (tens of microseconds against >= 0.1 s)."""
    from oracle import oracle as o
    from dusk_zerocaf_b200 import synth
    o.build()
    key = n_sample
    if key not in _cpu_cache:
        _cpu_cache.clear()
        a = synth.synth_fe(1, 0, n_sample)
        b_ = synth.synth_fe(2, 0, n_sample)
        prod = np.ones((n_sample, 5), dtype=np.uint64)           # ones: every page is written now
        sq = np.ones((n_sample, 5), dtype=np.uint64)
        _cpu_cache[key] = (a, b_, prod, sq)
    a, b_, prod, sq = _cpu_cache[key]
    o.fe_mul_square_batch(a[:4096], b_[:4096], threads=threads)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.fe_mul_square_batch(a, b_, threads=threads, out=(prod, sq))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2.0 * n_sample / best, best


def cpu_secondary_rates(threads):
    """Reference-equivalent CPU rates for configs 3-5 on bounded samples: point adds (2^16), strict
    scalar-muls (2^10), and the naive MSM the reference would run (sum of double_and_add, 2^10 points)."""
    from oracle import oracle as o
    from dusk_zerocaf_b200 import synth
    o.build()
    n_sm = 1 << 10
    base = np.tile(synth.BASEPOINT, (n_sm, 1))
    r = synth.synth_scalar(100, 0, n_sm)
    t0 = time.perf_counter()
    P = o.pt_scalar_mul_batch(base, r, threads=threads)
    t_sm = time.perf_counter() - t0
    n_add = 1 << 16
    PP = np.tile(P, (n_add // n_sm, 1))
    QQ = np.roll(PP, 7, axis=0)
    t0 = time.perf_counter()
    o.pt_add_batch(PP, QQ, threads=threads)
    t_add = time.perf_counter() - t0
    s = synth.synth_scalar(102, 0, n_sm)
    t0 = time.perf_counter()
    o.msm_naive(P, s, threads=threads)
    t_msm = time.perf_counter() - t0
    n_c = 1 << 12
    t0 = time.perf_counter()
    o.ris_compress_batch(P[:n_sm].repeat(n_c // n_sm, axis=0), threads=threads)
    t_c = time.perf_counter() - t0
    return {"ristretto_compress_per_s": n_c / t_c, "ristretto_compress_sample": n_c,
            "point_adds_per_s": n_add / t_add, "point_add_sample": n_add,
            "scalar_muls_per_s": n_sm / t_sm, "scalar_mul_sample": n_sm,
            "msm_2p20_seconds_extrapolated": t_msm * ((1 << 20) / n_sm), "msm_sample_points": n_sm,
            "note": "oracle port (1:1 restatement of edwards.rs:102-120, 465-489), linear extrapolation for the MSM"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = N_FIELD                                           # the whole config-2 batch: same workload as the B200 arm
    for _ in range(max(args.warmup, 1)):
        cpu_field_rate(n_sample, cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_field_rate(n_sample, cores)
        times.append(dt)
    total = sum(times)
    value = 2.0 * n_sample * args.steps / total
    v1, _ = cpu_field_rate(1 << 21, 1)                           # the reference itself is single-threaded: 1-thread figure beside it
    sample = (f"all 2^24 element pairs per step (mul+square), {cores} pthreads, oracle/zerocaf_oracle.c; inputs and outputs "
              f"allocated and touched before the timed region")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (radix-2^52 limbs, u128 products)", "data": "synthetic",
        "config": {"workload": "config 2: batched 2^24 FieldElement mul+square+reduce per step, reference AoS [u64;5] layout in and out",
                   "n_pairs_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_thread_value": v1, "single_thread_sample": "2^21 pairs, 1 thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import dusk_zerocaf_b200 as zc
    from dusk_zerocaf_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: zerocaf_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NUMA placement before anything pins host memory: this rank's CPUs = the ones next to its GPU (no-op on a one-node VM)
    from dusk_zerocaf_b200 import hostmem
    placement = hostmem.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    ctx = zc.Context(local, stream=stream.cuda_stream)
    L = ctx._L

    def timed(fn, steps, warmup):
        """W warm-up calls, then K calls bracketed by barrier+sync, CUDA events on the launching stream, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        stream.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), ctx.launches - l0

    def dev_u64(arr):
        return torch.from_numpy(arr.view(np.int64)).to(dev)

    # ---- config 2 inputs (each rank its own batch: stream ids differ per rank) -----------------------------------
    n = N_FIELD
    pin = lambda shape: torch.empty(shape, dtype=torch.int64).pin_memory()
    ha, hb, hp, hs = pin((n, 5)), pin((n, 5)), pin((n, 5)), pin((n, 5))
    synth.synth_fe(1 + 16 * rank, 0, n, out=ha.numpy().view(np.uint64))
    synth.synth_fe(2 + 16 * rank, 0, n, out=hb.numpy().view(np.uint64))
    da, db = ha.to(dev), hb.to(dev)
    dp, ds = torch.empty_like(da), torch.empty_like(da)
    torch.cuda.synchronize()

    def step_dev():
        ctx.check(L.zc_fe_mul_square_batch_dev(ctx._h, da.data_ptr(), db.data_ptr(), dp.data_ptr(), ds.data_ptr(), n))

    def step_e2e():
        ctx.check(L.zc_fe_mul_square_batch(ctx._h, ha.data_ptr(), hb.data_ptr(), hp.data_ptr(), hs.data_ptr(), n))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)                      # nvidia-smi needs a moment before its first sample
    ms, launches = timed(step_dev, args.steps, args.warmup)
    # the timed region of K launches lasts milliseconds, shorter than one nvidia-smi sample: keep the SAME kernel
    # running back to back for ~1.5 s right after it so the sampler sees the clocks under this load
    t_end = time.time() + (0.0 if args.no_sustain else 1.5)
    while time.time() < t_end:
        for _ in range(50):
            step_dev()
        stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_over"] = "timed region + 1.5 s of back-to-back launches of the same kernel immediately after it"
    value = 2.0 * n * world * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps

    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e, _ = timed(step_e2e, e2e_steps, max(args.warmup, 3))
    e2e_value = 2.0 * n * world * e2e_steps / (ms_e2e * 1e-3)
    # device result of the e2e path must equal the resident path's (same inputs)
    same = bool(torch.equal(hp.to(dev), dp) and torch.equal(hs.to(dev), ds))

    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_PAIR * n / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("fe_mul_square_kernel_bytes_per_launch")
    except Exception:
        pass

    # ---- integer-pipe roofline (configs 3-5 are multiplier-bound, so report both rooflines) ------------------
    # peak = 32 wide multiply-adds / clk / SM: an IMAD.WIDE.U32[.X] of a carry chain occupies the fmaheavy pipe for 4 cycles per
    # warp instruction (tools/ubench/pipes2-4 + montbench, profiles/r02_ubench_pipes.txt; cross-checked by ncu: the config-2
    # kernel issues 196 of them per warp in 1049 cycles with sm__pipe_fmaheavy_cycles_active = 78.9 %, i.e. 4.2 pipe cycles
    # each, profiles/r02_fe_mul_square_ncu.csv).  Scaled to the SM clock sampled under THIS load (the board sits on its 1 kW
    # power cap).  A carry-free mad.wide with all operands in the reuse cache issues at 61 / clk / SM in pipes.cu; a
    # carry-free product built on it (9 x 28-bit limbs) measured slower than the chained one (profiles/r02_pipes5_*.txt).
    WIDE_PER_CLK_SM, WIDE_PLAIN_UBENCH = 32.0, 61.0
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sm_count = 148

    def roofline_int(work_per_unit, units_per_s, note):
        peak = WIDE_PER_CLK_SM * sm_count * sm_mhz * 1e6
        ach = work_per_unit * units_per_s
        return {"work_per_unit": work_per_unit, "achieved": ach, "peak": peak, "unit": "wide-mults/s (32x32+64 -> 64)", "frac": ach / peak,
                "peak_per_clk_per_sm": WIDE_PER_CLK_SM, "plain_mad_wide_ubench_per_clk_per_sm": WIDE_PLAIN_UBENCH, "sm_mhz": sm_mhz, "note": note}

    # mul-only and square-only rates (N / t_mul beside the fused 2N / t)
    dm = torch.empty_like(da)
    ms_mul, _ = timed(lambda: ctx.check(L.zc_fe_mul_batch_dev(ctx._h, da.data_ptr(), db.data_ptr(), dm.data_ptr(), n)), max(5, args.steps // 2), 3)
    ms_sq, _ = timed(lambda: ctx.check(L.zc_fe_square_batch_dev(ctx._h, da.data_ptr(), dm.data_ptr(), n)), max(5, args.steps // 2), 3)
    kk = max(5, args.steps // 2)
    mul_only = {"mul_per_s": n * world * kk / (ms_mul * 1e-3), "ms_mul": ms_mul / kk, "square_per_s": n * world * kk / (ms_sq * 1e-3),
                "ms_square": ms_sq / kk, "hbm_frac_mul": 96.0 * n / (ms_mul / kk * 1e-3) / 1e9 / peak,
                "roofline_int_mul": roofline_int(112, n * kk / (ms_mul * 1e-3), "112 wide multiplies per Mul: 64 (8x8 words) + 32 + 16 (two folds with 2^K = -c)")}
    del dm
    # ---- the same workload on the 32-byte wire format (to_bytes / from_bytes, field.rs:563-631): resident and end to end ----
    pa, pb_ = torch.empty((n, 32), dtype=torch.uint8, device=dev), torch.empty((n, 32), dtype=torch.uint8, device=dev)
    pp, pq = torch.empty_like(pa), torch.empty_like(pa)
    ctx.check(L.zc_fe_to_bytes_batch_dev(ctx._h, da.data_ptr(), pa.data_ptr(), n))
    ctx.check(L.zc_fe_to_bytes_batch_dev(ctx._h, db.data_ptr(), pb_.data_ptr(), n))
    ms_pk, _ = timed(lambda: ctx.check(L.zc_fe_mul_square_batch_packed_dev(ctx._h, pa.data_ptr(), pb_.data_ptr(), pp.data_ptr(), pq.data_ptr(), n)), args.steps, 3)
    chk = torch.empty_like(da)
    ctx.check(L.zc_fe_from_bytes_batch_dev(ctx._h, pp.data_ptr(), chk.data_ptr(), n))
    ctx.sync()
    packed_same = bool(torch.equal(chk, dp))
    hpa, hpb = pa.cpu().pin_memory(), pb_.cpu().pin_memory()
    hpp, hpq = torch.empty((n, 32), dtype=torch.uint8).pin_memory(), torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    ms_pk_e2e, _ = timed(lambda: ctx.check(L.zc_fe_mul_square_batch_packed(ctx._h, hpa.data_ptr(), hpb.data_ptr(), hpp.data_ptr(), hpq.data_ptr(), n)), e2e_steps, 3)
    traffic_pk = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_pk = json.load(f).get("fe_mul_square_packed_kernel_bytes_per_launch")
    except Exception:
        pass
    ach_pk = BYTES_PER_PAIR * n / (ms_pk / args.steps * 1e-3) / 1e9
    packed = {"entry_point": "zc_fe_mul_square_batch_packed(_dev): 32-byte little-endian encodings in and out (the reference's to_bytes format)",
              "value": 2.0 * n * world * args.steps / (ms_pk * 1e-3), "unit": UNIT, "ms_per_step": ms_pk / args.steps,
              "roofline": {"bound": "hbm", "achieved": ach_pk, "peak": peak, "unit": "GB/s", "frac": ach_pk / peak, "traffic": traffic_pk},
              "e2e": {"value": 2.0 * n * world * e2e_steps / (ms_pk_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_pk_e2e / e2e_steps,
                      "h2d_bytes_per_step": 2 * n * 32 * world, "d2h_bytes_per_step": 2 * n * 32 * world},
              "matches_limb_layout_results": packed_same and bool(torch.equal(hpp.to(dev), pp))}
    del pa, pb_, pp, pq, chk, hpa, hpb, hpp, hpq
    rl_int = roofline_int(196, n / (kernel_ms * 1e-3), "196 wide multiplies per pair: Mul 112 + Square 84 (36 + 32 + 16); per GPU")

    del ha, hb, hp, hs, dp, ds, da, db
    extra = {}
    if not args.skip_extra:
        # ---- points for configs 3-5: P_i = [r_i]B from our own fixed-base kernel (Z != 1) ----------------
        def make_points(stream_id, count):
            sc = dev_u64(synth.synth_scalar(stream_id, 0, count))
            out = torch.empty((count, 20), dtype=torch.int64, device=dev)
            ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, sc.data_ptr(), out.data_ptr(), count))
            ctx.sync()
            return out

        k3 = max(3, args.steps // 2)
        P = make_points(100, N_POINT)
        Q = make_points(101, N_POINT)
        O = torch.empty_like(P)
        ms_add, _ = timed(lambda: ctx.check(L.zc_point_add_batch_dev(ctx._h, P.data_ptr(), Q.data_ptr(), O.data_ptr(), N_POINT)), k3, 3)
        ms_dbl, _ = timed(lambda: ctx.check(L.zc_point_double_batch_dev(ctx._h, P.data_ptr(), O.data_ptr(), N_POINT)), k3, 3)
        extra["config3_point_add"] = {
            "n": N_POINT, "adds_per_s": N_POINT * world * k3 / (ms_add * 1e-3), "ms": ms_add / k3,
            "doubles_per_s": N_POINT * world * k3 / (ms_dbl * 1e-3), "ms_double": ms_dbl / k3,
            "hbm_frac_add": 384.0 * N_POINT / (ms_add / k3 * 1e-3) / 1e9 / peak,
            "roofline_int": roofline_int(12 * 104, N_POINT * k3 / (ms_add * 1e-3), "12 Montgomery products of 96 wide + 8 low multiplies per limb-exact Add; per GPU")}
        # ---- config 4 ----------------------------------------------------------------------------------------------------
        S = dev_u64(synth.synth_scalar(102, 0, N_SMUL))
        P4, O4 = P[:N_SMUL].contiguous(), torch.empty((N_SMUL, 20), dtype=torch.int64, device=dev)
        ms_strict, _ = timed(lambda: ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, P4.data_ptr(), S.data_ptr(), O4.data_ptr(), N_SMUL, 0)), 2, 1)
        ms_fast, _ = timed(lambda: ctx.check(L.zc_point_scalar_mul_batch_dev(ctx._h, P4.data_ptr(), S.data_ptr(), O4.data_ptr(), N_SMUL, 1)), 2, 1)
        ms_fixed, _ = timed(lambda: ctx.check(L.zc_basepoint_mul_batch_dev(ctx._h, S.data_ptr(), O4.data_ptr(), N_SMUL)), 3, 1)
        extra["config4_scalar_mul"] = {
            "fixed_base_per_s": N_SMUL * world * 3 / (ms_fixed * 1e-3), "fixed_base_ms": ms_fixed / 3,
            "n": N_SMUL, "strict_per_s": N_SMUL * world * 2 / (ms_strict * 1e-3), "strict_ms": ms_strict / 2,
            "fast_per_s": N_SMUL * world * 2 / (ms_fast * 1e-3), "fast_ms": ms_fast / 2,
            "implied_point_adds_per_s_strict": 374.0 * N_SMUL * world * 2 / (ms_strict * 1e-3),
            "roofline_int_strict": roofline_int(374 * 12 * 104, N_SMUL * 2 / (ms_strict * 1e-3),
                                                "reference schedule: 249 doublings + ~125 additions, each the 12-product limb-exact Add; per GPU"),
            "roofline_int_fast": roofline_int((63 * 4 + 189 * 3 + 63 * 7 + 8 + 64) * 104 + 252 * 4 * 76, N_SMUL * 2 / (ms_fast * 1e-3),
                                              "4-bit signed windows: 252 dedicated doublings (4 squarings of 68 + 8 multiplies each, plus 4 products when T is "
                                              "needed -- 63 of them -- else 3) + 64 cached additions (7 products, the last 8) + the 8-entry table (64); per GPU")}
        # ---- SURVEY.md 8f "next" rows (wire formats either side of the path, hash to group, scalar-vector ops): rates at 2^20 ---
        n8 = 1 << 20
        P8 = P[:n8]
        S2 = dev_u64(synth.synth_scalar(103, 0, n8))
        enc = torch.empty((n8, 32), dtype=torch.uint8, device=dev)
        pts8 = torch.empty((n8, 20), dtype=torch.int64, device=dev)
        ok8 = torch.empty(n8, dtype=torch.uint8, device=dev)
        xy8 = torch.empty((n8, 10), dtype=torch.int64, device=dev)
        fe8 = torch.empty((n8, 5), dtype=torch.int64, device=dev)
        ub8 = torch.randint(0, 256, (n8, 64), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        naf8 = torch.empty((n8, 256), dtype=torch.int8, device=dev)
        f8 = {}

        def rate8(tag, fn, k=3):
            ms_, _ = timed(fn, k, 1)
            f8[tag + "_per_s"] = n8 * world * k / (ms_ * 1e-3)
            f8[tag + "_ms"] = ms_ / k
        rate8("ristretto_compress", lambda: ctx.check(L.zc_ristretto_compress_batch_dev(ctx._h, P8.data_ptr(), enc.data_ptr(), n8)))
        rate8("ristretto_decompress", lambda: ctx.check(L.zc_ristretto_decompress_batch_dev(ctx._h, enc.data_ptr(), pts8.data_ptr(), ok8.data_ptr(), n8)))
        ctx.sync()
        f8["decompress_all_valid"] = bool(ok8.all().item())
        rate8("point_to_affine", lambda: ctx.check(L.zc_point_to_affine_batch_dev(ctx._h, P8.data_ptr(), xy8.data_ptr(), n8)))
        rate8("point_is_valid", lambda: ctx.check(L.zc_point_is_valid_batch_dev(ctx._h, P8.data_ptr(), ok8.data_ptr(), n8)))
        rate8("fe_invert", lambda: ctx.check(L.zc_fe_invert_batch_dev(ctx._h, S.data_ptr(), fe8.data_ptr(), n8)))
        rate8("fe_pow", lambda: ctx.check(L.zc_fe_pow_batch_dev(ctx._h, S.data_ptr(), S2.data_ptr(), fe8.data_ptr(), n8)))
        rate8("elligator", lambda: ctx.check(L.zc_ristretto_elligator_batch_dev(ctx._h, S.data_ptr(), pts8.data_ptr(), n8)))
        rate8("from_uniform_bytes", lambda: ctx.check(L.zc_ristretto_from_uniform_bytes_batch_dev(ctx._h, ub8.data_ptr(), pts8.data_ptr(), n8)))
        rate8("scalar_mul_mod_l", lambda: ctx.check(L.zc_scalar_mul_batch_dev(ctx._h, S.data_ptr(), S2.data_ptr(), fe8.data_ptr(), n8)), 10)
        rate8("scalar_window_naf5", lambda: ctx.check(L.zc_scalar_window_naf_batch_dev(ctx._h, S.data_ptr(), 5, naf8.data_ptr(), n8)))
        # integer work of the inversion: Montgomery's trick over 8 elements per thread -- 3 products per element + one a^(p-2) chain
        # (251 dedicated squarings, popcount - 1 products) and two scaling products per 8 elements
        e_inv = (1 << 252) + 27742317777372353535851937790883648493 - 2
        inv_work = 3 * 104 + ((e_inv.bit_length() - 1) * 76 + (bin(e_inv).count("1") - 1 + 2) * 104) / 8.0
        f8["roofline_int_fe_invert"] = roofline_int(inv_work, f8["fe_invert_per_s"] / world,
                                                    "batched inversion, 8 elements per thread: 3 products (96 + 8 multiplies) per element + (251 squarings of 68 + 8 "
                                                    "and 66 products) / 8; per GPU")
        f8["note"] = ("SURVEY.md 8f ranks 1-4 through their _dev entry points, 2^20 elements each, outputs bit-identical to the reference "
                      "methods (tests/test_gpu_parity.py); decompress runs on the encodings compress just produced")
        extra["section8f_next_rows"] = f8
        del enc, pts8, ok8, xy8, fe8, ub8, naf8, S2
        # ---- config 5: MSM, strong scaling over the N ranks (bucket-window sharding + one exchange) ----------------------
        # Three modes, same scalars: plain (arbitrary points, operand pass inside the call), prepared generators (handle,
        # Z = 1 cached operands), fixed-base tables (handle, pre-scaled rows).  At N > 1 every mode is ALSO timed on one GPU
        # in the same run (rank 0's device, while the others wait), so the speed-ups in msm_scaling are same-run, same-box.
        out_pt = torch.zeros(20, dtype=torch.int64, device=dev)
        km = max(5, args.steps // 2)
        single = {}

        def time_single(tag, fn):
            # every rank times its own single-GPU MSM (they all hold the same inputs); max over ranks like everything else
            ms_, _ = timed(fn, km, 3)
            single[tag] = ms_ / km

        plain1 = lambda: ctx.check(L.zc_msm_dev(ctx._h, P4.data_ptr(), S.data_ptr(), N_MSM, MSM_WINDOW, out_pt.data_ptr()))
        gens_prep = ctx.msm_generators(P4.data_ptr(), N_MSM, zc.GEN_PREPARED)
        time_single("plain", plain1)
        time_single("prepared", lambda: gens_prep.msm(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW))
        ctx.sync()
        single_pt = out_pt.clone()
        t0 = time.perf_counter()
        gens_fb1 = ctx.msm_generators(P4.data_ptr(), N_MSM, zc.GEN_FIXED_BASE, MSM_WINDOW, 0, 1)
        fb1_prepare_ms = (time.perf_counter() - t0) * 1e3
        time_single("fixed_base", lambda: gens_fb1.msm(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW))
        fb1_bytes = gens_fb1.device_bytes
        gens_fb1.close()
        ms_msm_nccl = ms_msm_bcast = None
        oracle_ok = None
        if world > 1:
            def bcast(b):
                obj = [b]
                dist.broadcast_object_list(obj, src=0)
                return obj[0]

            def allgather(b):
                objs = [None] * world
                dist.all_gather_object(objs, b)
                return objs
            ctx.init_nccl(rank, world, bcast)
            plainN = lambda: ctx.check(L.zc_msm_sharded_dev(ctx._h, P4.data_ptr(), S.data_ptr(), N_MSM, MSM_WINDOW, out_pt.data_ptr()))
            ms_msm_nccl, _ = timed(plainN, 5, 3)                # exchange = ncclAllGather + fold kernel
            ctx.sync()
            nccl_pt = out_pt.clone()
            # ---- oracle parity of BOTH exchange paths, outside any timed region: n = 2^14 points, every rank
            # runs the collective, rank 0 compares with the CPU oracle's naive MSM (fold of double_and_add, ristretto.rs:166-176)
            n_chk = 1 << 14
            chk = {}
            chk_pt = torch.zeros(20, dtype=torch.int64, device=dev)
            ctx.check(L.zc_msm_sharded_dev(ctx._h, P4.data_ptr(), S.data_ptr(), n_chk, MSM_WINDOW, chk_pt.data_ptr()))
            ctx.sync()
            chk["nccl"] = chk_pt.cpu().numpy().view(np.uint64).copy()
            ctx.init_peer_mailboxes(rank, world, allgather)     # from here on: NVLink peer stores + flags + fold, one kernel
            dist.barrier()
            ctx.check(L.zc_msm_sharded_dev(ctx._h, P4.data_ptr(), S.data_ptr(), n_chk, MSM_WINDOW, chk_pt.data_ptr()))
            ctx.sync()
            chk["peer_mailboxes"] = chk_pt.cpu().numpy().view(np.uint64).copy()
            if rank == 0:
                from oracle import oracle as o          # the checker, outside the timed regions
                o.build()
                want = o.msm_naive(P4[:n_chk].cpu().numpy().view(np.uint64), S[:n_chk].cpu().numpy().view(np.uint64), threads=os.cpu_count() or 1)
                oracle_ok = {k_: bool(o.pt_eq(v, want)) for k_, v in chk.items()}
                oracle_ok["n_points"] = n_chk
            msm_plain = plainN
            msm_prep = lambda: gens_prep.msm_sharded(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW)
        else:
            msm_plain = plain1
            msm_prep = lambda: gens_prep.msm(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW)
        ms_msm_raw, _ = timed(msm_plain, km, 3)                 # arbitrary points: operand pass inside every call
        ms_msm, l_msm = timed(msm_prep, km, 3)                  # fixed generators: prepared once, the timed call only sees new scalars
        ctx.sync()
        prep_pt = out_pt.clone()
        identical, same_elem = True, True
        eq = torch.zeros(2, dtype=torch.uint8, device=dev)
        if world > 1:
            g = [torch.zeros_like(out_pt) for _ in range(world)]
            dist.all_gather(g, prep_pt)
            identical = all(bool(torch.equal(g[0], x)) for x in g)
            both = torch.stack([single_pt, nccl_pt])
            ref2 = torch.stack([prep_pt, prep_pt])
            ctx.check(L.zc_ristretto_eq_batch_dev(ctx._h, both.data_ptr(), ref2.data_ptr(), eq.data_ptr(), 2))
            ctx.sync()
            same_elem = bool(eq.all().item())
            # second figure: the scalars are new for every MSM and live on rank 0 -- one NCCL broadcast of the
            # 40 MiB limb array (32 MiB of information) in front of every call, on the MSM's stream
            try:
                def msm_bcast_fn():
                    with torch.cuda.stream(stream):
                        dist.broadcast(S, src=0)
                    msm_prep()
                ms_msm_bcast, _ = timed(msm_bcast_fn, km, 2)
            except Exception as e:                              # never let the secondary figure take the bench line down
                print(f"[bench] scalar-broadcast figure skipped: {e}", file=sys.stderr)
                ms_msm_bcast = None
        # fixed generators, memory traded for time: pre-scaled per-window tables -> one merged bucket set per rank, no doubling chain
        t0 = time.perf_counter()
        gens_fb = ctx.msm_generators(P4.data_ptr(), N_MSM, zc.GEN_FIXED_BASE, MSM_WINDOW, rank, world)
        fb_prepare_ms = (time.perf_counter() - t0) * 1e3
        msm_fb = (lambda: gens_fb.msm_sharded(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW)) if world > 1 else \
                 (lambda: gens_fb.msm(S.data_ptr(), out_pt.data_ptr(), window_bits=MSM_WINDOW))
        ms_msm_fb, l_msm_fb = timed(msm_fb, km, 3)
        ctx.sync()
        fb_pt = out_pt.clone()
        eqf = torch.zeros(1, dtype=torch.uint8, device=dev)
        ctx.check(L.zc_ristretto_eq_batch_dev(ctx._h, fb_pt.data_ptr(), single_pt.data_ptr(), eqf.data_ptr(), 1))
        ctx.sync()
        fb_same = bool(eqf.all().item())
        fb_identical = True
        if world > 1:
            g = [torch.zeros_like(fb_pt) for _ in range(world)]
            dist.all_gather(g, fb_pt)
            fb_identical = all(bool(torch.equal(g[0], x)) for x in g)
        fb_bytes = gens_fb.device_bytes
        gens_fb.close()
        gens_prep.close()
        # the MSM's end-to-end figure: host pointers (pinned), 200 MiB of H2D inside the timed region (rank 0 only, one GPU)
        msm_e2e_ms = None
        if world == 1:
            hP, hS = P4.cpu().pin_memory(), S.cpu().pin_memory()
            hO = torch.zeros(20, dtype=torch.int64).pin_memory()
            ms_h, _ = timed(lambda: ctx.check(L.zc_msm(ctx._h, hP.data_ptr(), hS.data_ptr(), N_MSM, MSM_WINDOW, hO.data_ptr())), 3, 2)
            msm_e2e_ms = ms_h / 3
            del hP, hS
        # integer work of one MSM: 16 x 2^20 bucket additions of 7 (prepared) / 8 (plain) Montgomery products +
        # the reductions and 240 doublings (~2 % more)
        adds = 16.0 * N_MSM
        modes = {"plain": ms_msm_raw / km, "prepared": ms_msm / km, "fixed_base": ms_msm_fb / km}
        best1 = min(single.values())
        bestN = min(modes.values())
        extra["msm_scaling"] = {
            "n_points": N_MSM, "window_bits": MSM_WINDOW, "n_gpus": world, "scaling": "strong",
            "ms_single_gpu_same_run": single, "ms_n_gpus": modes,
            "speedup_per_mode": {k_: single[k_] / modes[k_] for k_ in modes},
            "speedup_best_vs_best": best1 / bestN, "best_single_gpu_mode": min(single, key=single.get), "best_n_gpu_mode": min(modes, key=modes.get),
            "msm_per_s_best": 1e3 / bestN,
            "note": "plain = arbitrary points (operand pass in the call); prepared / fixed_base = generator handles (zc_msm_generators). "
                    "Single-GPU times are measured in this same run on every rank's own GPU (max over ranks)."}
        extra["config5_msm_fixed_base_tables"] = {
            "n_points": N_MSM, "window_bits": MSM_WINDOW, "n_gpus": world, "msm_per_s": km / (ms_msm_fb * 1e-3),
            "ms_per_msm": ms_msm_fb / km, "scaling": "strong", "launches_per_msm": l_msm_fb / km,
            "table_bytes_per_rank": fb_bytes, "prepare_ms_once": fb_prepare_ms,
            "single_gpu_table_bytes": fb1_bytes, "single_gpu_prepare_ms_once": fb1_prepare_ms,
            "matches_plain_single_gpu_msm": fb_same, "all_ranks_identical_bits": fb_identical}
        extra["config5_msm"] = {
            "n_points": N_MSM, "window_bits": MSM_WINDOW, "n_gpus": world, "msm_per_s": km / (ms_msm * 1e-3),
            "ms_per_msm": ms_msm / km, "scaling": "strong", "launches_per_msm": l_msm / km,
            "points_prepared": True, "ms_per_msm_unprepared": ms_msm_raw / km,
            "all_ranks_identical_bits": identical, "matches_single_gpu_and_nccl_path": same_elem,
            "matches_oracle": oracle_ok,
            "exchange": "NVLink peer-memory mailboxes: peer stores + flags + tree fold in one kernel" if world > 1 else None,
            "ms_per_msm_nccl_exchange": (ms_msm_nccl / 5) if world > 1 else None,
            "ms_per_msm_incl_scalar_broadcast": (ms_msm_bcast / km) if ms_msm_bcast else None,
            "ms_per_msm_e2e_host_pointers": msm_e2e_ms,
            "sharding": "bucket-window (w mod N), one exchange of the 160-B partial points + fixed-order fold" if world > 1 else "single GPU",
            "hbm_frac": 160.0 * N_MSM / (ms_msm / km * 1e-3) / 1e9 / peak,
            "roofline_int": roofline_int(adds * 7 * 104 * 1.02, km / (ms_msm * 1e-3),
                                         "whole job: 16 x 2^20 mixed additions x 7 Montgomery products x 104 multiplies (+2 % reductions / chain); "
                                         "achieved is the aggregate over the N GPUs, peak is ONE GPU's -- divide frac by n_gpus for per-GPU utilisation")}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        ns = 1 << 23
        v, dt = cpu_field_rate(ns, cores, repeats=3)
        v1, dt1 = cpu_field_rate(1 << 21, 1, repeats=1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2^23 of the 2^24 pairs (mul+square), best of 3, {cores} pthreads, oracle/zerocaf_oracle.c "
                         f"(1:1 restatement of field.rs:250-262,302-315; no Rust toolchain here)",
               "single_thread_value": v1, "configs_3_to_5": cpu_secondary_rates(cores)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8 x 32-bit words per residue, IMAD.WIDE carry chains); ABI u64x5 radix-2^52", "data": "synthetic",
            "config": {"workload": "config 2: batched 2^24 FieldElement mul+square+reduce per GPU, reference AoS [u64;5] layout in and out",
                       "n_pairs_per_step_per_gpu": n, "e2e_path": "zc_fe_mul_square_batch (host pointers, pinned; chunked H2D / kernel / D2H pipeline inside the library)", "l2_policy": "inputs larger than L2 (1.34 GB read + 1.34 GB written per step)",
                       "parallelism": f"replicated element-wise x{world}, no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic bytes = 128 B per pair (32-B canonical encodings); the kernel moves 160 B per pair in the reference's 40-B limb layout"},
            "roofline_int": rl_int,
            "mul_only": mul_only,
            "packed_wire_format": packed,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 40 * world, "d2h_bytes_per_step": 2 * n * 40 * world,
                    "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps, "matches_resident_path": same},
            "gpu_launches": launches,
            "host_placement": placement,
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-extra", action="store_true", help="only the config-2 line (used under ncu)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-sustain", action="store_true", help="skip the 1.5 s clock-sampling loop (ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" and not args.skip_extra else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
