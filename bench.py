#!/usr/bin/env python
"""Benchmark of the hot path: SNP-gene GxC interaction tests per second of `run_interaction`
(BASELINE.json metric; workload = configs[2]: n = 100k cells, 1,000 donors, k = 20 contexts, low-rank hK
(q = 50 -> m = 1,020), 10k SNPs per GPU).

    python bench.py --gpus N --steps K --warmup W                 # this implementation (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU restatement of the reference path

One step = one whole `run_interaction` job over the rank's SNP shard: constructor set-up (Gram,
11 eigendecompositions) + rotation + 11 REML fits/SNP + score statistic + Davies/Liu p-values.
`value` times it with every input already resident in HBM; `e2e` times the public API call with host
buffers (host->device copies of y, E, W, hK and the genotype matrix, device->host read of the results inside
the timed region).  SNPs shard across ranks with no data-path collective except the final all-gather of the
5 per-SNP outputs ("scaling": "weak": every rank scans its own 10k SNPs of the same gene).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SNP-gene GxC tests/sec (run_interaction, n=100k cells, k=20)"
UNIT = "tests/s"
# measured on this pool's B200 (profiles/r01_dmma_probe.txt): DMMA.8x8x4 issue-rate peak; cuBLAS DGEMM reaches 35.4
FP64_TENSOR_PEAK_TFLOPS = 37.1
# dram__bytes_read.sum + dram__bytes_write.sum of the rotation launch at the default workload, from one ncu capture
# (profiles/r01_ncu_rotation_traffic_benchsize.csv): 465.7 GB read + 1.8 GB written per 10k-SNP launch, against 27 GB of
# algorithmic bytes (G 8 GB + pre-expanded basis 17.2 GB + output 1.7 GB) -- operand panels are re-read through L2 by the
# 13 272 CTAs; 395 GB/s = 6 % of the HBM peak, the kernel is bound by the FP64 tensor pipe (99 % active).
ROTATION_DRAM_BYTES_DEFAULT_WORKLOAD = 465728855296 + 1766195200
# the same for one launch of oz_mma_kernel at the default workload (profiles/r01_ncu_oz_mma_kernel.txt): 320.7 GB read + 6.9 GB
# written against 19.9 GB of algorithmic bytes (17.2 GB digit planes + 1.0 GB int8 dosages + 1.7 GB output) -- the 148 persistent
# CTAs re-stream their operand panels through L2 (hit rate 64 %); 3.0 TB/s = 36 % of the HBM peak, the kernel is bound by the int8
# tensor pipe under the power cap and by L2 -> SM bandwidth (47.7 B/clk/SM)
OZ_MMA_DRAM_BYTES_DEFAULT_WORKLOAD = 320706872000 + 6889898000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload overrides (defaults = BASELINE configs[2]); used by the tests to run a tiny instance
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--donors", type=int, default=1000)
    ap.add_argument("--contexts", type=int, default=20)
    ap.add_argument("--hk-rank", type=int, default=50)
    ap.add_argument("--snps", type=int, default=10000, help="SNPs per GPU")
    ap.add_argument("--cpu-sample-snps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-donor-level", action="store_true")
    ap.add_argument("--no-fp64-route", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"run_interaction n={a.cells} cells, {a.donors} donors, k={a.contexts}, low-rank hK q={a.hk_rank} "
            f"(m={a.contexts * (1 + a.hk_rank)}), {a.snps} SNPs per GPU")


# ------------------------------------------------------------------------------------------------
# synthetic gene: host part (small arrays) + genotype shard
# ------------------------------------------------------------------------------------------------
def make_gene(a, seed=0):
    """y, W, E, hK of one gene (identical on every rank) -- recipe of cellregmap_b200/synth.py at scale."""
    rng = np.random.default_rng(seed)
    n, d, k, q = a.cells, a.donors, a.contexts, a.hk_rank
    donor = np.sort(rng.integers(0, d, n))
    E = rng.standard_normal((n, k))
    E = (E - E.mean(0)) / E.std(0) / np.sqrt(k)
    A = rng.standard_normal((d, q)) / np.sqrt(q)
    hK = np.ascontiguousarray(A[donor])
    W = np.ones((n, 1))

    def mom(v):
        v = v - v.mean()
        return v / v.std()

    y = np.full(n, 0.3)
    y += np.sqrt(0.25) * mom(E @ rng.standard_normal(k))
    y += np.sqrt(0.25) * mom(((hK @ rng.standard_normal((q, k))) * E).sum(1))
    y += np.sqrt(0.45) * rng.standard_normal(n)
    return {"y": y, "W": W, "E": np.ascontiguousarray(E), "hK": hK, "donor": donor, "rng_state": seed}


def donor_genotypes(a, rank, seed=0):
    rng = np.random.default_rng(1000 + 17 * rank + seed)
    maf = rng.uniform(0.05, 0.45, a.snps)
    Gd = rng.binomial(2, maf, size=(a.donors, a.snps)).astype(np.float64)
    for j in np.where(Gd.std(0) == 0)[0]:
        Gd[rng.integers(0, a.donors), j] += 1.0
    return Gd


def add_causal_effects(gene, Gd_rank0, a):
    """persistent effects of SNPs 5, 6 and GxC effects of SNPs 10, 11 of rank 0's shard (reference test recipe)."""
    rng = np.random.default_rng(99)
    donor, E, y = gene["donor"], gene["E"], gene["y"]

    def norm(v):
        v = v - v.mean()
        s = v.std()
        return v / s if s > 0 else v

    if a.snps > 11:
        g = np.stack([Gd_rank0[:, j][donor] for j in (5, 6, 10, 11)], 1)
        y += np.sqrt(0.03) * norm(norm(g[:, 0]) * rng.standard_normal() + norm(g[:, 1]) * rng.standard_normal())
        y += np.sqrt(0.02) * norm(norm(g[:, 2]) * (E @ rng.standard_normal(E.shape[1])) + norm(g[:, 3]) * (E @ rng.standard_normal(E.shape[1])))
    return gene


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(device_index):
    """Bind this process to the CPUs NVML reports as local to the GPU, so that the pinned host buffers of the e2e arm (first touch)
    and the driver calls stay on the GPU's socket: a buffer on the far socket halves the host-to-device rate.  Best effort."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:       # NVML missing, no affinity information, restricted container ...
        return None


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines, self.t_mark = gpu_index, None, [], 0.0

    def mark(self):
        """Samples from now on count (the sampler is started before the warm-up so that the start-up of nvidia-smi, which can
        stall other CUDA work for hundreds of milliseconds, does not fall into the timed region)."""
        self.t_mark = time.time()

    def start(self):
        if os.environ.get("CRM_BENCH_NO_SAMPLER") == "1":       # A/B switch for the sampler's own footprint
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, smax, reasons = [], [], set()
        for t_line, line in self.lines:
            if t_line < self.t_mark:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(a, gene, Gd, n_snps, repeats=1):
    """Times the reference algorithm (oracle port: per-(SNP, rho1) LMM construction + rotation + Brent,
    structured projection, Davies) on the first `n_snps` SNPs.  Set-up (11 economic decompositions) is built
    once, outside the timed region, via the Gram route (qs_method="gram") so that the run stays bounded; per-SNP
    cost in the reference does not depend on the number of SNPs, so tests/s = SNPs / scan seconds."""
    from oracle import crm_port
    from threadpoolctl import threadpool_limits
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1; the reference's only parallelism is the BLAS under numpy, so give it every core
    with threadpool_limits(limits=threads):
        t0 = time.time()
        Ls = crm_port.get_L_values(gene["hK"], gene["E"])
        model = crm_port.CellRegMapOracle(y=gene["y"], E=gene["E"], W=gene["W"], E1=gene["E"], Ls=Ls, qs_method="gram")
        setup_s = time.time() - t0
        G = np.ascontiguousarray(Gd[:, :n_snps][gene["donor"]])
        times = []
        for _ in range(repeats):
            t0 = time.time()
            model.scan_interaction(G)
            times.append(time.time() - t0)
    return {"scan_s": times, "setup_s": setup_s, "cores": threads, "snps": n_snps}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gene = make_gene(a)
    Gd = donor_genotypes(a, 0)
    gene = add_causal_effects(gene, Gd, a)
    ns = max(1, min(a.cpu_sample_snps, a.snps))
    res = cpu_reference_sample(a, gene, Gd, ns, repeats=a.warmup + a.steps)
    timed = res["scan_s"][a.warmup:]
    ms = 1e3 * float(np.mean(timed))
    value = ns / (ms / 1e3)
    sample = (f"first {ns} SNPs of the same gene per step, scan only; set-up ({res['setup_s']:.1f} s, Gram-route decompositions) "
              f"built once outside the timed region; numpy/BLAS threads = {res['cores']}")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    full_affinity = os.sched_getaffinity(0)
    local_cpus = None if os.environ.get("CRM_BENCH_NO_PIN") == "1" else pin_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"       # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import cellregmap_b200 as crm
    from cellregmap_b200 import _cellregmap as api
    from cellregmap_b200 import _lib

    lib = _lib.load()
    gene = make_gene(a)
    Gd0 = donor_genotypes(a, 0)
    gene = add_causal_effects(gene, Gd0, a)
    Gd = Gd0 if rank == 0 else donor_genotypes(a, rank)
    p = a.snps
    # device-resident inputs
    y_d = torch.from_numpy(gene["y"]).to(dev)
    W_d = torch.from_numpy(gene["W"]).to(dev)
    E_d = torch.from_numpy(gene["E"]).to(dev)
    hK_d = torch.from_numpy(gene["hK"]).to(dev)
    donor_d = torch.from_numpy(gene["donor"]).to(dev)
    G_d = torch.from_numpy(Gd).to(dev)[donor_d].contiguous()          # (n, p) float64 dosages, expanded donor -> cell
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = torch.empty((world * 5, p), dtype=torch.float64, device=dev) if world > 1 else None

    debug = bool(os.environ.get("CRM_BENCH_DEBUG"))

    def step_device():
        t0 = time.time()
        model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
        if debug:
            torch.cuda.synchronize(); t1 = time.time()
        out = model._scan_interaction_device(G_d)
        if debug:
            torch.cuda.synchronize(); t2 = time.time()
            print(f"[debug] rank {rank}: model {1e3 * (t1 - t0):.1f} ms, scan {1e3 * (t2 - t1):.1f} ms", file=sys.stderr)
        res = torch.stack([out["pv"], out["rho1"], out["e2"], out["g2"], out["eps2"]])
        if world > 1:   # the path's one exchange step: all-gather of the 5 per-SNP outputs
            dist.all_gather_into_tensor(gathered, res)
            return gathered
        return res

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        barrier()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, r

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(a.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler.mark()
    api.PROFILE.update(on=True, rot_ms=0.0, rot_flops=0.0, rot_launches=0, int8_ms=0.0, int8_ops=0.0, int8_launches=0)
    launches0 = lib.crm_launch_count()
    ms_total, wall, res = timed(step_device, a.steps)
    launches = lib.crm_launch_count() - launches0
    api.PROFILE["on"] = False
    clocks = sampler.stop()
    ms_per_step = max(ms_total, wall * 1e3) / a.steps     # device events and host wall clock bracket the same region
    value = world * p / (ms_per_step / 1e3)
    rot_ms, rot_flops, rot_launches = api.PROFILE["rot_ms"], api.PROFILE["rot_flops"], api.PROFILE["rot_launches"]
    achieved = rot_flops / (rot_ms * 1e-3) / 1e12 if rot_ms > 0 else None
    int8_ms, int8_ops, int8_launches = api.PROFILE["int8_ms"], api.PROFILE["int8_ops"], api.PROFILE["int8_launches"]
    pv = res[0]
    top = torch.argsort(pv)[:4].tolist()
    alg_flop = 2.0 * a.cells * a.contexts * (1 + a.hk_rank) * (1 + a.contexts)
    default_workload = (a.cells, a.donors, a.contexts, a.hk_rank, a.snps) == (100000, 1000, 20, 50, 10000)
    if int8_launches > 0:
        # Integer dosages: the rotation ran as the exact int8 split.  Its dominant kernel is oz_mma_kernel (hand-written
        # tcgen05.mma kind::i8 + TMA + TMEM, fp64 recombination of the digit planes fused into the epilogue); its roofline is
        # int8 TOP/s against 2 x the measured dense bf16 peak.  CRM_INT8_GEMM=lt swaps in cuBLASLt + a recombination kernel.
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        int8_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
        int8_tops = int8_ops / (int8_ms * 1e-3) / 1e12
        fused = os.environ.get("CRM_INT8_GEMM") != "lt" and os.environ.get("CRM_INT8_MMA") != "2cta"
        roofline = {"bound": "tensor", "achieved": int8_tops, "peak": int8_peak, "unit": "TOP/s (int8)", "frac": int8_tops / int8_peak,
                    "traffic": OZ_MMA_DRAM_BYTES_DEFAULT_WORKLOAD if default_workload and fused else None,
                    "traffic_unit": "bytes per launch (ncu, profiles/r01_ncu_oz_mma_kernel.txt)",
                    "kernel": ("cuBLASLt int8 GEMM + oz_combine_kernel" if os.environ.get("CRM_INT8_GEMM") == "lt" else
                               "oz_mma_kernel (hand-written tcgen05.mma kind::i8 + TMA, TMEM accumulators, fused fp64 recombination)") +
                              ": 8 digit planes of [Hx|Hx.E_j] against int8 dosages, the exact int8 split of the rotation",
                    "peak_source": ("2 x bf16_tflops of MEASURED_PEAKS.json" if peaks else "2 x 1.59 PFLOP/s fallback of B200_PROFILING.md") +
                                   " (int8 dense = 2 x bf16 dense on B200; both the measured bf16 figure and this kernel are limited by the "
                                   "power cap, so frac can exceed 1: against the nominal 4.5 POP/s the kernel reaches 0.73-0.76)",
                    "launches": int(int8_launches), "ms_per_launch": int8_ms / max(1, int8_launches), "share_of_step": int8_ms / ms_total if ms_total else None,
                    "rotation_fp64_equivalent": {"achieved": achieved, "unit": "TFLOP/s", "algorithmic_flop_per_test": alg_flop,
                                                 "ms_per_launch": rot_ms / max(1, rot_launches), "share_of_step": rot_ms / ms_total if ms_total else None,
                                                 "vs_fp64_tensor_peak": (achieved / FP64_TENSOR_PEAK_TFLOPS) if achieved else None,
                                                 "note": "whole rotation (int8 conversion + digit planes + GEMM + fp64 recombination) in algorithmic fp64 flop; "
                                                         "above the 37.1 TFLOP/s FP64 tensor peak because the work runs on the int8 tensor pipe"}}
    else:
        roofline = {"bound": "tensor", "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": (achieved / FP64_TENSOR_PEAK_TFLOPS) if achieved else None,
                    "traffic": ROTATION_DRAM_BYTES_DEFAULT_WORKLOAD if default_workload and api.PROFILE.get("pre_expanded_basis") else None,
                    "traffic_unit": "bytes per launch (ncu, profiles/r01_ncu_rotation_traffic_benchsize.csv)",
                    "kernel": ("crm_gemm_kernel<PLAIN> on the pre-expanded basis [Hx|Hx.E_j] (rotation of [g, g.E] onto [H|y|W])"
                               if api.PROFILE.get("pre_expanded_basis") else "crm_gemm_kernel<EXPAND> (rotation, Hadamard on the fly)"),
                    "algorithmic_flop_per_test": alg_flop, "launches": int(rot_launches), "ms_per_launch": rot_ms / max(1, rot_launches),
                    "share_of_step": rot_ms / ms_total if ms_total else None,
                    "peak_source": "FP64 DMMA issue-rate peak measured on this pool (profiles/r01_dmma_probe.txt); cuBLAS DGEMM 35.4"}

    # ---- the same job on the fp64 tensor-core route (hand-written DMMA kernel; what non-integer genotypes get) ----
    fp64_route = None
    if int8_launches > 0 and not a.no_fp64_route:
        os.environ["CRM_ROTATION"] = "dmma"
        try:
            step_device()
            api.PROFILE.update(on=True, rot_ms=0.0, rot_flops=0.0, rot_launches=0, int8_ms=0.0, int8_ops=0.0, int8_launches=0)
            ms_f, wall_f, res_f = timed(step_device, a.steps)
            api.PROFILE["on"] = False
        finally:
            del os.environ["CRM_ROTATION"]
        ms_f = max(ms_f, wall_f * 1e3) / a.steps
        ach_f = api.PROFILE["rot_flops"] / (api.PROFILE["rot_ms"] * 1e-3) / 1e12 if api.PROFILE["rot_ms"] > 0 else None
        pos = (res[0] > 0) & (res_f[0] > 0)
        fp64_route = {"value": world * p / (ms_f / 1e3), "unit": UNIT, "ms_per_step": ms_f,
                      "max_abs_dlog10p_vs_int8_split": float((torch.log10(res_f[0][pos]) - torch.log10(res[0][pos])).abs().max()),
                      "roofline": {"bound": "tensor", "achieved": ach_f, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                                   "frac": (ach_f / FP64_TENSOR_PEAK_TFLOPS) if ach_f else None,
                                   "traffic": ROTATION_DRAM_BYTES_DEFAULT_WORKLOAD if default_workload else None,
                                   "traffic_unit": "bytes per launch (ncu, profiles/r01_ncu_rotation_traffic_benchsize.csv)",
                                   "kernel": "crm_gemm_kernel<PLAIN> (hand-written DMMA + TMA) on the pre-expanded basis [Hx|Hx.E_j]",
                                   "launches": int(api.PROFILE["rot_launches"]), "ms_per_launch": api.PROFILE["rot_ms"] / max(1, api.PROFILE["rot_launches"]),
                                   "peak_source": "FP64 DMMA issue-rate peak measured on this pool (profiles/r01_dmma_probe.txt); cuBLAS DGEMM 35.4"}}

    # ---- extension: donor-level genotype ingress (same job, G given as donors x SNPs + donor index) ----
    donor_level = None
    if not a.no_donor_level:
        Gdon_d = torch.from_numpy(Gd).to(dev)

        def step_donor():
            model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
            out = model._scan_interaction_device(Gdon_d, donor_index=donor_d)
            return torch.stack([out["pv"], out["rho1"], out["e2"], out["g2"], out["eps2"]])

        step_donor()
        ms_d, wall_d, res_d = timed(step_donor, a.steps)
        ms_d = max(ms_d, wall_d * 1e3) / a.steps
        dl = None
        if world == 1:      # compare where p > 0 (the strongest simulated hits underflow to exactly 0 in both paths)
            pos = (res[0] > 0) & (res_d[0] > 0)
            assert bool(((res[0] > 0) == (res_d[0] > 0)).all())
            dl = float((torch.log10(res_d[0][pos]) - torch.log10(res[0][pos])).abs().max())
        donor_level = {"value": world * p / (ms_d / 1e3), "unit": UNIT, "ms_per_step": ms_d, "max_abs_dlog10p_vs_expanded": dl,
                       "note": "same job with genotypes passed as a (donors x SNPs) matrix + donor index (keyword-only extension of the "
                               "reference API); the per-SNP contraction runs over donors instead of cells; not the headline value"}

    # ---- extension: many genes over one data set -- the model is kept, only the phenotype changes per step ----
    shared_setup = None
    if not a.no_donor_level:
        keep = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
        y_alt = [y_d + 0.01 * i * torch.sin(torch.arange(a.cells, device=dev, dtype=torch.float64)) for i in range(1, 3)]
        counter = {"i": 0}

        def step_shared(G_in, **kw):
            counter["i"] += 1
            keep.set_phenotype(y_alt[counter["i"] % 2])
            out = keep._scan_interaction_device(G_in, **kw)
            return torch.stack([out["pv"], out["rho1"], out["e2"], out["g2"], out["eps2"]])

        shared_setup = {"note": "model object kept across steps, CellRegMap.set_phenotype(y) per step (extension for scans of many genes "
                                "over the same cells); not the headline value", "unit": UNIT}
        for name, fn in (("expanded_genotypes", lambda: step_shared(G_d)), ("donor_level_genotypes", lambda: step_shared(Gdon_d, donor_index=donor_d))):
            fn()
            ms_s, wall_s, _ = timed(fn, a.steps)
            ms_s = max(ms_s, wall_s * 1e3) / a.steps
            shared_setup[name] = {"value": world * p / (ms_s / 1e3), "ms_per_step": ms_s}
        del keep

    # ---- e2e through the public API with (pinned) host buffers ----
    e2e = None
    if not a.no_e2e:
        try:
            G_h = torch.empty((a.cells, p), dtype=torch.float64, pin_memory=True)
        except RuntimeError:        # the host cannot page-lock 8 GB per rank: pageable memory (block-streamed transfer)
            G_h = torch.empty((a.cells, p), dtype=torch.float64)
        G_h.copy_(G_d)
        y_h, W_h, E_h, hK_h = (torch.from_numpy(gene[key]).pin_memory() for key in ("y", "W", "E", "hK"))
        torch.cuda.synchronize()

        def step_host():
            pv_h, info_h = crm.run_interaction(y_h, E_h, G_h, W=W_h, hK=hK_h)     # numpy results = D2H inside
            if world > 1:
                res_h = torch.from_numpy(np.stack([pv_h, info_h["rho1"], info_h["e2"], info_h["g2"], info_h["eps2"]])).to(dev)
                dist.all_gather_into_tensor(gathered, res_h)
            return pv_h

        step_host()
        ms_e, wall_e, pv_h = timed(step_host, a.steps)
        ms_e = max(ms_e, wall_e * 1e3) / a.steps
        h2d = G_h.numel() * 8 + sum(t.numel() * 8 for t in (y_h, W_h, E_h, hK_h))
        e2e = {"value": world * p / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(5 * p * 8),
               "ms_per_step": ms_e}
        assert np.array_equal(pv_h, res[5 * rank if world > 1 else 0].cpu().numpy()), "host and device paths disagree"
        # the same call with the dosages stored as int8 on the host (8x less PCIe traffic); reported beside the float64 headline
        try:
            G_h8 = torch.empty((a.cells, p), dtype=torch.int8, pin_memory=True)
        except RuntimeError:
            G_h8 = torch.empty((a.cells, p), dtype=torch.int8)
        G_h8.copy_(G_d)
        del G_h

        def step_host8():
            pv8, info8 = crm.run_interaction(y_h, E_h, G_h8, W=W_h, hK=hK_h)
            if world > 1:
                res8 = torch.from_numpy(np.stack([pv8, info8["rho1"], info8["e2"], info8["g2"], info8["eps2"]])).to(dev)
                dist.all_gather_into_tensor(gathered, res8)
            return pv8

        step_host8()
        ms_8, wall_8, pv_8 = timed(step_host8, a.steps)
        ms_8 = max(ms_8, wall_8 * 1e3) / a.steps
        assert np.array_equal(pv_8, pv_h), "int8 and float64 host genotypes disagree"
        e2e["int8_host_genotypes"] = {"value": world * p / (ms_8 / 1e3), "unit": UNIT, "ms_per_step": ms_8,
                                      "h2d_bytes_per_step": int(G_h8.numel() + sum(t.numel() * 8 for t in (y_h, W_h, E_h, hK_h))),
                                      "note": "same API call with the genotype matrix stored as int8 on the host; not the headline e2e"}
        del G_h8

    # ---- CPU baseline (rank 0, N = 1 only) ----
    os.sched_setaffinity(0, full_affinity)      # the CPU arm uses every host core again
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        ns = max(1, min(a.cpu_sample_snps, p))
        r = cpu_reference_sample(a, gene, Gd, ns)
        v = ns / r["scan_s"][0]
        cpu = {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"first {ns} SNPs of the same gene, scan only ({r['scan_s'][0]:.1f} s); set-up ({r['setup_s']:.1f} s) excluded"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(a), "snps_per_gpu": p, "l2": "inputs (8 GB genotypes, 0.8 GB basis) far larger than L2",
                           "host_affinity": ("%d CPUs local to the GPU (NVML)" % len(local_cpus)) if local_cpus else "unchanged",
                           "step": "constructor set-up + scan of the rank's SNP shard + all-gather of results"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "fp64_route": fp64_route,
                "cpu_baseline": cpu, "donor_level_ingress": donor_level, "shared_setup": shared_setup,
                "top_hits": top}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
