#!/usr/bin/env python
"""bench.py — headline benchmark of qdsp_b200 (BASELINE.json configs[1]).

Workload: the fused NCO xlator -> polyphase resampler (2.4 MS/s -> 48 kS/s, 401 taps, I=1, D=50) -> FM
quadrature demod chain on 2^28 synthetic cf32 samples per GPU, cut into the reference's run() blocks of
819 200 samples. Metric: Msamples/s of INPUT cf32 consumed.

  python bench.py [--gpus N] [--steps K] [--warmup W]        our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                        the reference's own CPU chain, same metric

One JSON line on stdout (rank 0). `value` = whole-job throughput with the input resident in HBM;
`e2e` = the same chain through the C-ABI host-buffer entry point (pinned host in/out, H2D + D2H timed);
`roofline` = the fused kernel's algorithmic bytes / its own CUDA-event duration vs the measured HBM peak;
`cpu_baseline` = the unmodified reference headers (oracle/_ref) timed on this box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 1 << 28          # per GPU (config 2)
BLOCK = 819_200              # reference-sized run() block with count*I % D == 0
FS, FC, FM, DEV = 2_400_000, 250_000, 1_000, 5e3
OUT_SR, BW = 48e3, 48e3
ALG_BYTES_PER_SAMPLE = 8.0 + 4.0 / 50.0   # cf32 in + f32 audio out per input sample (cfg 2)
ALG_FLOP_PER_SAMPLE = 6 + 401 * 4 / 50 + 0.4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# =================================================================================================
# reference arm / cpu_baseline: the UNMODIFIED reference chain on host cores
# =================================================================================================
_CPU_INPUTS = {}


def cpu_reference_run(sample: int, replicas: int):
    """Time `replicas` independent copies of the reference chain (VFO -> FloatFMDemod, 3 worker threads
    each, as the reference schedules them) over disjoint shards of a `sample`-long synthetic stream.
    Returns (Msamples/s, cores, kind, description)."""
    from oracle import loader
    from qdsp_b200 import synth

    per = (sample // replicas // BLOCK) * BLOCK
    per = max(per, BLOCK)
    key = (per, replicas)
    if _CPU_INPUTS.get("key") != key:   # synthetic input generated once, outside every timed region
        _CPU_INPUTS["key"] = key
        _CPU_INPUTS["xs"] = [synth.cfg2_input(r * per, per) for r in range(replicas)]
    xs = _CPU_INPUTS["xs"]
    if loader.have_ref("fast"):
        R = loader.ref("fast")
        kind, what = "reference", "unmodified reference headers (VFO + FloatFMDemod, 3 threads/chain) + VOLK shim, -O3 -march=x86-64-v3"

        def run(x):
            return R.vfo_fm(float(FC), float(FS), OUT_SR, BW, DEV, x, BLOCK, timing=True)[2]
    else:
        P = loader.port()
        kind, what = "port", "oracle/port.c restatement (single thread per chain)"

        def run(x):
            t0 = time.perf_counter()
            P.vfo_fm(float(FC), float(FS), OUT_SR, BW, DEV, x, BLOCK)
            return time.perf_counter() - t0
    results = [None] * replicas

    def worker(i):
        results[i] = run(xs[i])

    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(replicas)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    cores = replicas * (3 if kind == "reference" else 1)
    return per * replicas / wall / 1e6, cores, kind, f"{replicas} x {per} samples of the config-2 stream in {BLOCK}-sample blocks; {what}"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    replicas = max(1, min(ncpu // 3, 16))
    sample = replicas * 8 * BLOCK       # ~6.5 M samples per chain per step
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, kind, desc = cpu_reference_run(sample, replicas)
        if i >= args.warmup:
            vals.append(v)
    v = float(np.mean(vals))
    n_per_step = (sample // replicas // BLOCK) * BLOCK * replicas
    line = {
        "impl": "reference", "metric": "Msamples/s cf32 through xlate-resample-demod chain", "value": v, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_per_step / v / 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: NCO xlator + polyphase resampler 2.4MS/s->48kS/s (401 taps, I=1, D=50) + FM demod; "
                               "bounded CPU sample per step", "samples_per_step": n_per_step, "block": BLOCK},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =================================================================================================
# our arm
# =================================================================================================
def ours(args):
    import torch
    import torch.distributed as dist

    from qdsp_b200 import blocks as B, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; qdsp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    L = lib.load()
    lib.check(L.qdsp_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.samples
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    # ---- synthetic input, generated on the device (each rank: its own stream = an independent channel)
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    lib.check(L.qdsp_synth_fm_cf32(x.data_ptr(), rank * n, n, FS, FC, FM, DEV, 0.5, 0.005, 2, sp))
    chain = B.VFOFM(float(FC), float(FS), OUT_SR, BW, DEV)
    n_out = chain.out_count(n, BLOCK)
    audio = torch.empty(n_out + 64, dtype=torch.float32, device=dev)
    L.qdsp_vfofm_enable_timing(chain.h, 1)

    def step():
        m = L.qdsp_vfofm_process(chain.h, x.data_ptr(), audio.data_ptr(), None, n, None, 0, BLOCK, None, sp)
        if m != n_out:
            raise SystemExit(f"process returned {m}, expected {n_out}: {lib.last_error()}")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.qdsp_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # one CUDA event pair per launch of the fused kernel, recorded on the launching stream inside the timed region
    lib.check(L.qdsp_vfofm_enable_timing_ring(chain.h, args.steps), "enable_timing_ring")
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.qdsp_launch_count() - launches0
    # dominant-kernel duration: mean of the K event pairs recorded around the fused kernel during the timed region
    import ctypes as _C
    nrec = _C.c_int(0)
    kernel_ms = [float(L.qdsp_vfofm_kernel_ms_mean(chain.h, _C.byref(nrec)))]
    assert nrec.value == args.steps, (nrec.value, args.steps)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end: pinned host in -> C-ABI host entry point -> pinned host out ---------------------
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(n, dtype=torch.complex64, pin_memory=True)
        xh.copy_(x)
        yh = torch.empty(n_out + 64, dtype=torch.float32, pin_memory=True)
        torch.cuda.synchronize()
        chain2 = B.VFOFM(float(FC), float(FS), OUT_SR, BW, DEV)

        def e2e_step():
            m = L.qdsp_vfofm_process_host(chain2.h, xh.data_ptr(), yh.data_ptr(), n, BLOCK, sp)
            if m != n_out:
                raise SystemExit(f"process_host returned {m}, expected {n_out}: {lib.last_error()}")

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()          # synchronises the stream internally: result is in host memory on return
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * args.e2e_steps / float(tt.item()) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n_out * 4, "steps": args.e2e_steps,
               "checksum": float(yh[:n_out].double().abs().sum())}
        del xh, yh

    # the sampler ran through the device-resident steps, the per-kernel timing pass and the end-to-end steps
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    k_ms = float(np.mean(kernel_ms))
    achieved = ALG_BYTES_PER_SAMPLE * n / (k_ms * 1e-3) / 1e9
    # DRAM bytes per launch of the fused kernel at the default workload, from the committed ncu --set full capture
    # (profiles/r01_fused_vfofm_final_ncu_full.txt: dram__bytes_read.sum 2.264114 GB + dram__bytes_write.sum 26.07 MB)
    traffic = 2.264114e9 + 26.067968e6 if n == N_SAMPLES else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "qdsp::decim_sup_kernel<9,50,5,ROT,DEMOD> (fused xlate+resample+demod)",
            "kernel_ms": k_ms, "alg_bytes_per_launch": ALG_BYTES_PER_SAMPLE * n,
            "fp32_tflops": ALG_FLOP_PER_SAMPLE * n / (k_ms * 1e-3) / 1e12}
    cpu = None
    if not args.no_cpu:
        ncpu = os.cpu_count() or 1
        v, cores, kind, desc = cpu_reference_run(8 * BLOCK * 1, 1)
        # run again with all cores the reference's threading model can use
        replicas = max(1, min(ncpu // 3, 16))
        v_all, cores_all, _, desc_all = cpu_reference_run(replicas * 8 * BLOCK, replicas)
        cpu = {"value": v_all, "unit": "Msamples/s", "cores": cores_all, "kind": kind, "sample": desc_all,
               "single_chain": {"value": v, "cores": cores, "sample": desc}, "host_cpus": ncpu}
    line = {
        "metric": "Msamples/s cf32 through xlate-resample-demod chain", "value": value, "unit": "Msamples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: NCO xlator + polyphase resampler 2.4MS/s->48kS/s (401 taps, I=1, D=50) + FM demod, fused",
                   "samples_per_gpu": n, "block": BLOCK, "parallelism": f"independent streams x{world} (no collective)",
                   "l2": "input (2 GiB/GPU) is larger than L2: no flush needed"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=N_SAMPLES)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
