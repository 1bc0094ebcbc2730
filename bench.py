#!/usr/bin/env python
"""Benchmark of the hot path: SNP-gene GxC interaction tests per second of `run_interaction`
(BASELINE.json metric; default workload = configs[2]: n = 100k cells, 1,000 donors, k = 20 contexts, low-rank hK
(q = 50 -> m = 1,020), 10k SNPs sharded over the GPUs).

    python bench.py --gpus N --steps K --warmup W                   # this implementation (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's own code on the host cores
    python bench.py --config 1|2|4|5                                # the other BASELINE.json configs (1 GPU; bench lines for BASELINE.md)

One step = one whole job: constructor set-up (Gram, 11 eigendecompositions) + rotation + 11 REML fits/SNP + score statistic +
Davies/Liu p-values for every SNP.  `value` times it with every input already resident in HBM; `e2e` times the public API call
with plain (pageable) numpy arrays -- what a user of the reference passes -- host->device transfers and the device->host read of
the results inside the timed region.  N > 1 (`"scaling": "strong"`, configs[2] as written): the 10k SNPs of the one gene are sharded
over the ranks (cellregmap_b200.distributed.run_interaction_sharded: shared set-up, per-rank column blocks, one all-gather of the
5 per-SNP outputs); `weak_scaling` in the same line repeats the measurement with 10k SNPs per GPU.
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

UNIT = "tests/s"
# BASELINE.json configs: (entry point, cells, donors, contexts, hK rank, SNPs)
CONFIGS = {
    1: ("run_interaction", 500, 50, 10, 50, 100),
    2: ("run_interaction", 10000, 200, 20, 10, 2000),
    3: ("run_interaction", 100000, 1000, 20, 50, 10000),
    4: ("run_association", 50000, 1000, 20, 50, 10000),
    5: ("estimate_betas", 50000, 1000, 20, 50, 1000),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the default workload, from ncu --set full captures
# of the same command (profiles/): kernel name -> bytes
NCU_TRAFFIC = {}
try:
    NCU_TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
except (OSError, ValueError):
    pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: --snps sharded over the ranks (configs[2]) or --snps per rank")
    # workload overrides (defaults come from --config); used by the tests to run a tiny instance
    ap.add_argument("--cells", type=int)
    ap.add_argument("--donors", type=int)
    ap.add_argument("--contexts", type=int)
    ap.add_argument("--hk-rank", type=int)
    ap.add_argument("--snps", type=int, help="SNPs of the job (strong scaling: in total)")
    ap.add_argument("--cpu-sample-snps", type=int, default=0, help="SNPs of the CPU sample (0: sized for ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary arms (fp64 route, donor-level ingress, shared set-up, weak scaling)")
    a = ap.parse_args()
    entry, cells, donors, contexts, q, snps = CONFIGS[a.config]
    a.entry = entry
    a.cells = a.cells or cells
    a.donors = a.donors or donors
    a.contexts = a.contexts or contexts
    a.hk_rank = a.hk_rank or q
    a.snps = a.snps or snps
    return a


def metric_name(a):
    if a.entry == "run_interaction":
        return f"SNP-gene GxC tests/sec (run_interaction, n={a.cells // 1000}k cells, k={a.contexts})" if a.cells >= 1000 else \
            f"SNP-gene GxC tests/sec (run_interaction, n={a.cells} cells, k={a.contexts})"
    if a.entry == "run_association":
        return f"SNP-gene association tests/sec (run_association, n={a.cells // 1000}k cells, k={a.contexts})"
    return f"SNPs/sec with effect sizes (estimate_betas, n={a.cells // 1000}k cells, k={a.contexts})"


def workload_name(a):
    return (f"configs[{a.config - 1}]: {a.entry} n={a.cells} cells, {a.donors} donors, k={a.contexts}, low-rank hK q={a.hk_rank} "
            f"(m={a.contexts * (1 + a.hk_rank) if a.entry != 'run_association' else a.contexts + a.hk_rank}), {a.snps} SNPs")


# ------------------------------------------------------------------------------------------------
# synthetic gene: host part (small arrays) + genotypes
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


def donor_genotypes(a, part, snps, seed=0):
    """donors x snps dosages; `part` selects an independent stream (weak scaling: one per rank)."""
    rng = np.random.default_rng(1000 + 17 * part + seed)
    maf = rng.uniform(0.05, 0.45, snps)
    Gd = rng.binomial(2, maf, size=(a.donors, snps)).astype(np.float64)
    for j in np.where(Gd.std(0) == 0)[0]:
        Gd[rng.integers(0, a.donors), j] += 1.0
    return Gd


def add_causal_effects(gene, Gd, a):
    """persistent effects of SNPs 5, 6 and GxC effects of SNPs 10, 11 (reference test recipe)."""
    rng = np.random.default_rng(99)
    donor, E, y = gene["donor"], gene["E"], gene["y"]

    def norm(v):
        v = v - v.mean()
        s = v.std()
        return v / s if s > 0 else v

    if Gd.shape[1] > 11:
        g = np.stack([Gd[:, j][donor] for j in (5, 6, 10, 11)], 1)
        y += np.sqrt(0.03) * norm(norm(g[:, 0]) * rng.standard_normal() + norm(g[:, 1]) * rng.standard_normal())
        y += np.sqrt(0.02) * norm(norm(g[:, 2]) * (E @ rng.standard_normal(E.shape[1])) + norm(g[:, 3]) * (E @ rng.standard_normal(E.shape[1])))
    return gene


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(device_index, local_world=1, local_rank=0):
    """Bind this process to the CPUs NVML reports as local to the GPU (a pinned buffer on the far socket halves the host-to-device
    rate); with several ranks per box each takes its own slice of them, so that the host feeders of the ranks do not share cores.
    Best effort; returns the CPU list."""
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
        cpus = sorted(cpus)
        if cpus and local_world > 1:
            share = max(1, len(cpus) // local_world)
            mine = cpus[local_rank * share:(local_rank + 1) * share] if (local_rank + 1) * share <= len(cpus) else cpus
            cpus = mine or cpus
        if cpus:
            os.sched_setaffinity(0, set(cpus))
        return cpus
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
# CPU arm: the reference's own source (oracle/_ref, over the dependency stand-ins of oracle/) or the oracle port, bounded sample
# ------------------------------------------------------------------------------------------------
def default_cpu_sample(a):
    """SNPs for roughly 10-30 s of CPU work on a 16-core host (measured per-SNP costs of the reference path)."""
    if a.cpu_sample_snps > 0:
        return min(a.cpu_sample_snps, a.snps)
    per_config = {1: 100, 2: 400, 3: 40, 4: 100, 5: 2}
    if a.impl == "reference":      # the reference arm repeats the sample warmup + steps times: a third of it per step keeps the run within minutes
        per_config = {1: 100, 2: 160, 3: 16, 4: 40, 5: 1}
    return max(1, min(per_config[a.config], a.snps))


def cpu_reference_sample(a, gene, Gd, n_snps, repeats=1):
    """Times the reference algorithm on the first `n_snps` SNPs with all host threads.  The code that runs is the reference's own
    cellregmap/_cellregmap.py + _math.py (byte-compiled into oracle/_ref, imported over the stand-ins for glimix_core / numpy_sugar /
    chiscore) when it is present, else the oracle's restatement of it (same arithmetic, tests/test_reference_source.py).  The per-rho
    economic decompositions of the constructor take the Gram route (same Q0, S0; a thin SVD of 1e5 x 1020 eleven times would take
    minutes) and are built once, outside the timed region; per-SNP cost in the reference does not depend on the number of SNPs, so
    tests/s = SNPs / scan seconds."""
    from threadpoolctl import threadpool_limits
    from oracle import crm_port, ref_shims
    ref = ref_shims.load_reference(qs_method="gram")
    threads = os.cpu_count() or 1
    code = "reference source (oracle/_ref) over ported dependencies" if ref is not None else "oracle port"
    G = np.ascontiguousarray(Gd[:, :n_snps][gene["donor"]])
    y, W, E, hK = gene["y"], gene["W"], gene["E"], gene["hK"]
    # torchrun exports OMP_NUM_THREADS=1; the reference's only parallelism is the BLAS under numpy, so give it every core
    with threadpool_limits(limits=threads):
        t0 = time.time()
        if a.entry == "run_association":            # reference :471-500 (positional quirk: W and E swap roles)
            model = ref.CellRegMap(y, W, E, hK=hK) if ref is not None else crm_port.CellRegMapOracle(y, W, E, hK=hK, qs_method="gram")
            scan = lambda: model.scan_association(G)                       # noqa: E731
        else:
            get_L = sys.modules["cellregmap._cellregmap"].get_L_values if ref is not None else crm_port.get_L_values
            Ls = get_L(hK, E)
            model = ref.CellRegMap(y=y, E=E, W=W, E1=E, Ls=Ls) if ref is not None else crm_port.CellRegMapOracle(y=y, E=E, W=W, E1=E, Ls=Ls, qs_method="gram")
            if a.entry == "estimate_betas":         # reference :640-682 (per (SNP, rho1) decompositions inside the timed region)
                maf = crm_port.compute_maf(G)
                scan = lambda: model.predict_interaction(G, maf)           # noqa: E731
            else:
                scan = lambda: model.scan_interaction(G)                   # noqa: E731
        setup_s = time.time() - t0
        times = []
        for _ in range(repeats):
            t0 = time.time()
            scan()
            times.append(time.time() - t0)
    return {"scan_s": times, "setup_s": setup_s, "cores": threads, "snps": n_snps, "code": code}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("TQDM_DISABLE", "1")
    gene = make_gene(a)
    ns = default_cpu_sample(a)
    Gd = donor_genotypes(a, 0, a.snps)
    gene = add_causal_effects(gene, Gd, a)
    res = cpu_reference_sample(a, gene, Gd, ns, repeats=a.warmup + a.steps)
    timed = res["scan_s"][a.warmup:]
    ms = 1e3 * float(np.mean(timed))
    value = ns / (ms / 1e3)
    sample = (f"first {ns} SNPs of the same gene per step, scan only; set-up ({res['setup_s']:.1f} s, Gram-route decompositions) "
              f"built once outside the timed region; numpy/BLAS threads = {res['cores']}; code = {res['code']}")
    line = {"metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
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
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if a.entry != "run_interaction" and world > 1:
        raise SystemExit("--config 4 and 5 are single-GPU bench lines")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    full_affinity = os.sched_getaffinity(0)
    local_cpus = None if os.environ.get("CRM_BENCH_NO_PIN") == "1" else pin_to_gpu_numa_node(local_rank, local_world, local_rank)
    if local_cpus and "CRM_HOST_THREADS" not in os.environ:
        # feeder threads of this rank: its share of the box's cores, one of them left to the thread that drives the GPU when the share is small
        share = len(local_cpus)
        os.environ["CRM_HOST_THREADS"] = str(max(1, min(16, share - 1 if share <= 8 else share)))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"       # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    import cellregmap_b200 as crm
    from cellregmap_b200 import _cellregmap as api
    from cellregmap_b200 import _lib
    from cellregmap_b200 import distributed as crmd

    lib = _lib.load()
    gene = make_gene(a)
    p_total = a.snps                      # SNPs of the job the headline times
    Gd = donor_genotypes(a, 0, p_total)   # the same matrix on every rank (strong scaling: each rank scans its column block)
    gene = add_causal_effects(gene, Gd, a)
    y_d = torch.from_numpy(gene["y"]).to(dev)
    W_d = torch.from_numpy(gene["W"]).to(dev)
    E_d = torch.from_numpy(gene["E"]).to(dev)
    hK_d = torch.from_numpy(gene["hK"]).to(dev)
    donor_d = torch.from_numpy(gene["donor"]).to(dev)
    strong = world > 1 and a.scaling == "strong"
    lo, hi = crmd.shard_range(p_total, rank, world) if strong else (0, p_total)
    if world > 1 and not strong:          # weak scaling as the headline: every rank scans its own p_total SNPs
        Gd = donor_genotypes(a, rank, p_total) if rank else Gd
    # device-resident genotypes of this rank: (n, p_local) float64 dosages, expanded donor -> cell
    G_d = torch.from_numpy(np.ascontiguousarray(Gd[:, lo:hi])).to(dev)[donor_d].contiguous()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stack5(out):
        return torch.stack([out["pv"], out["rho1"], out["e2"], out["g2"], out["eps2"]])

    gathered_weak = torch.empty((world * 5, p_total), dtype=torch.float64, device=dev) if world > 1 else None

    # ---- the step, inputs resident in HBM ----
    def step_device():
        if a.entry == "run_association":
            return crm.run_association(y_d, W_d, E_d, G_d, hK=hK_d)[0]
        if a.entry == "estimate_betas":
            return crm.estimate_betas(y_d, W_d, E_d, G_d, maf=maf_d, hK=hK_d)[0]
        model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev, group=True if world > 1 else None)
        out = model._scan_interaction_device(G_d)
        if strong:      # the path's one exchange step: all-gather of the 5 per-SNP outputs
            return crmd.gather_results(stack5(out), p_total)
        if world > 1:
            dist.all_gather_into_tensor(gathered_weak, stack5(out))
            return gathered_weak
        return stack5(out)

    maf_d = None
    if a.entry == "estimate_betas":
        maf_d = torch.from_numpy(crm.compute_maf(Gd[:, lo:hi])).to(dev)

    debug = bool(os.environ.get("CRM_BENCH_DEBUG"))

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for i in range(steps):
            ts = time.time()
            r = fn()
            if debug:
                torch.cuda.synchronize()
                print(f"[debug] rank {rank} step {i}: {1e3 * (time.time() - ts):.1f} ms", file=sys.stderr, flush=True)
        e1.record()
        barrier()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return max(ms, wall * 1e3) / steps, r          # device events and host wall clock bracket the same region

    def warm_up(fn, at_least, at_most_extra=16):
        """`at_least` untimed steps, then more (untimed, bounded) until three consecutive steps agree within 5 %: the first models of a
        process grow the library's memory pool and the cuBLAS/cuSOLVER workspaces, which takes a varying number of steps to settle."""
        recent = []
        if os.environ.get("CRM_BENCH_FIXED_WARMUP") == "1":     # profiler runs: exactly the requested number of steps
            at_most_extra = 0
        i = -1
        for i in range(at_least + at_most_extra):
            torch.cuda.synchronize(); t0 = time.time()
            fn()
            torch.cuda.synchronize()
            recent = (recent + [time.time() - t0])[-3:]
            settled = len(recent) == 3 and max(recent) <= 1.05 * min(recent)
            if world > 1:       # every rank must take the same number of steps (collectives inside)
                flag = torch.tensor([1 if settled else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                settled = bool(int(flag[0]))
            if i + 1 >= at_least and settled:
                break
        return i + 1

    sampler = ClockSampler(local_rank)
    sampler.start()
    warm_steps = warm_up(step_device, a.warmup)
    torch.cuda.synchronize()
    sampler.mark()
    api.PROFILE.update(on=True, rot_ms=0.0, rot_flops=0.0, rot_flops_executed=0.0, rot_launches=0, int8_ms=0.0, int8_ops=0.0, int8_launches=0)
    launches0 = lib.crm_launch_count()
    ms_per_step, res = timed(step_device, a.steps)
    launches = lib.crm_launch_count() - launches0
    api.PROFILE["on"] = False
    clocks = sampler.stop()
    units = p_total if (world == 1 or strong) else world * p_total
    value = units / (ms_per_step / 1e3)
    ms_total = ms_per_step * a.steps

    # ---- roofline of the dominant kernel (interaction configs) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    fp64_peak = ctypes_double(lib.crm_fp64_tensor_peak)     # DMMA issue-rate peak measured on this device, now
    roofline, top = None, None
    rot_ms, rot_flops, rot_launches = api.PROFILE["rot_ms"], api.PROFILE["rot_flops"], api.PROFILE["rot_launches"]
    int8_ms, int8_ops, int8_launches = api.PROFILE["int8_ms"], api.PROFILE["int8_ops"], api.PROFILE["int8_launches"]
    default_workload = (a.config, a.cells, a.donors, a.contexts, a.hk_rank, a.snps, world) == (3, 100000, 1000, 20, 50, 10000, 1)
    if a.entry == "run_interaction":
        achieved = rot_flops / (rot_ms * 1e-3) / 1e12 if rot_ms > 0 else None
        executed = api.PROFILE["rot_flops_executed"] / (rot_ms * 1e-3) / 1e12 if rot_ms > 0 else None
        rows_full, rows_used = api.PROFILE.get("rotation_rows", (0, 0))
        rows_note = (f"{rows_used} of the {rows_full} columns of [Hx|Hx.E_j] are contracted: the background is L = (E V) x hK (get_L_values), so L.E_j = sum_l V_li "
                     "hK.E_l.E_j with triple products symmetric in (l, j); the rest is rebuilt by kr_expand_kernel (crm_set_background_factors)"
                     if rows_used and rows_used < rows_full else None)
        alg_flop = 2.0 * a.cells * a.contexts * (1 + a.hk_rank) * (1 + a.contexts)
        top = torch.argsort(res[0])[:4].tolist()
        if int8_launches > 0:
            # Integer dosages: the rotation ran as the exact int8 split.  Its dominant kernel is oz_mma_kernel (hand-written
            # tcgen05.mma kind::i8 + TMA + TMEM, fp64 recombination of the digit planes fused into the epilogue); its roofline is
            # int8 TOP/s against 2 x the measured dense bf16 peak.  CRM_INT8_GEMM=lt swaps in cuBLASLt + a recombination kernel.
            int8_peak = 2.0 * float(peaks.get("bf16_tflops", 1590.0))
            int8_tops = int8_ops / (int8_ms * 1e-3) / 1e12
            fused = os.environ.get("CRM_INT8_GEMM") != "lt" and os.environ.get("CRM_INT8_MMA") != "2cta"
            roofline = {"bound": "tensor", "achieved": int8_tops, "peak": int8_peak, "unit": "TOP/s (int8)", "frac": int8_tops / int8_peak,
                        "traffic": NCU_TRAFFIC.get("oz_mma_kernel") if default_workload and fused else None,
                        "traffic_unit": "bytes per launch (ncu --set full of this command, profiles/)",
                        "kernel": ("cuBLASLt int8 GEMM + oz_combine_kernel" if os.environ.get("CRM_INT8_GEMM") == "lt" else
                                   "oz_mma_kernel (hand-written tcgen05.mma kind::i8 + TMA, TMEM accumulators, fused fp64 recombination)") +
                                  ": digit planes of [Hx|Hx.E_j] (of its distinct columns when the background is structured, see contracted_columns) against int8 dosages, the exact int8 split of the rotation",
                        "peak_source": ("2 x bf16_tflops of MEASURED_PEAKS.json" if peaks else "2 x 1.59 PFLOP/s fallback of B200_PROFILING.md") +
                                       " (int8 dense = 2 x bf16 dense on B200; both the measured bf16 figure and this kernel are limited by the "
                                       "power cap, so frac can exceed 1: against the nominal 4.5 POP/s the kernel reaches 0.73-0.76)",
                        "launches": int(int8_launches), "ms_per_launch": int8_ms / max(1, int8_launches), "share_of_step": int8_ms / ms_total if ms_total else None,
                        # the path's algorithmic work is fp64 flop; reported against the FP64 tensor (DMMA) peak as well
                        "achieved_fp64_equivalent": achieved, "fp64_tensor_peak": fp64_peak, "frac_fp64_equivalent": (achieved / fp64_peak) if achieved and fp64_peak else None,
                        "fp64_equivalent_note": "whole rotation (int8 conversion + digit planes + contraction + fp64 recombination) in algorithmic fp64 flop "
                                                "(2 n m (1+k) per test) over its CUDA-event time, against the DMMA peak measured in this run; above 1 because "
                                                "the work runs on the int8 tensor pipe",
                        "algorithmic_flop_per_test": alg_flop, "rotation_ms_per_launch": rot_ms / max(1, rot_launches),
                        "rotation_share_of_step": rot_ms / ms_total if ms_total else None,
                        "ops_counted": "int8 operations the kernel executes (8 digit planes x 2 x cells x contracted columns x SNPs), not the "
                                       "algorithmic count of the full basis",
                        "contracted_columns": rows_note}
        else:
            roofline = {"bound": "tensor", "achieved": executed, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (executed / fp64_peak) if executed and fp64_peak else None,
                        "contracted_columns": rows_note,
                        "traffic": NCU_TRAFFIC.get("crm_gemm_kernel") if default_workload and api.PROFILE.get("pre_expanded_basis") else None,
                        "traffic_unit": "bytes per launch (ncu --set full of this command, profiles/)",
                        "kernel": ("crm_gemm_kernel<PLAIN> on the pre-expanded basis [Hx|Hx.E_j] (rotation of [g, g.E] onto [H|y|W])"
                                   if api.PROFILE.get("pre_expanded_basis") else "crm_gemm_kernel<EXPAND> (rotation, Hadamard on the fly)"),
                        "algorithmic_flop_per_test": alg_flop, "launches": int(rot_launches), "ms_per_launch": rot_ms / max(1, rot_launches),
                        "share_of_step": rot_ms / ms_total if ms_total else None, "frac_fp64_equivalent": (achieved / fp64_peak) if achieved and fp64_peak else None,
                        "peak_source": "FP64 DMMA issue-rate peak measured in this run (crm_fp64_tensor_peak)"}

    extras = not a.no_extras and a.entry == "run_interaction"
    # ---- the same job on the fp64 tensor-core route (hand-written DMMA kernel; what real-valued genotypes get) ----
    fp64_route = None
    if extras and world == 1 and int8_launches > 0:
        os.environ["CRM_ROTATION"] = "dmma"
        try:
            step_device()
            api.PROFILE.update(on=True, rot_ms=0.0, rot_flops=0.0, rot_flops_executed=0.0, rot_launches=0, int8_ms=0.0, int8_ops=0.0, int8_launches=0)
            ms_f, res_f = timed(step_device, max(2, a.steps // 2))
            api.PROFILE["on"] = False
        finally:
            del os.environ["CRM_ROTATION"]
        ach_f = api.PROFILE["rot_flops_executed"] / (api.PROFILE["rot_ms"] * 1e-3) / 1e12 if api.PROFILE["rot_ms"] > 0 else None
        alg_f = api.PROFILE["rot_flops"] / (api.PROFILE["rot_ms"] * 1e-3) / 1e12 if api.PROFILE["rot_ms"] > 0 else None
        pos = (res[0] > 0) & (res_f[0] > 0)
        fp64_route = {"value": p_total / (ms_f / 1e3), "unit": UNIT, "ms_per_step": ms_f,
                      "max_abs_dlog10p_vs_int8_split": float((torch.log10(res_f[0][pos]) - torch.log10(res[0][pos])).abs().max()),
                      "roofline": {"bound": "tensor", "achieved": ach_f, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (ach_f / fp64_peak) if ach_f and fp64_peak else None,
                                   "traffic": NCU_TRAFFIC.get("crm_gemm_kernel") if default_workload else None,
                                   "kernel": "crm_gemm_kernel<PLAIN> (hand-written DMMA + TMA) on the pre-expanded basis (its distinct columns, see "
                                             "roofline.contracted_columns) + kr_expand_kernel",
                                   "flop_counted": "executed (2 x cells x contracted columns x SNPs) over the CUDA-event time of the rotation incl. expansion",
                                   "algorithmic_tflops": alg_f, "frac_algorithmic": (alg_f / fp64_peak) if alg_f and fp64_peak else None,
                                   "launches": int(api.PROFILE["rot_launches"]), "ms_per_launch": api.PROFILE["rot_ms"] / max(1, api.PROFILE["rot_launches"]),
                                   "peak_source": "FP64 DMMA issue-rate peak measured in this run (crm_fp64_tensor_peak)"}}

    # ---- the same job with standardised genotype columns (the reference simulator's column_normalize, _simulate.py:50-54,339): affine
    #      images of integer dosages take the int8 contraction, anything else real-valued would take the fp64 route above ----
    standardised = None
    if extras and world == 1:
        G_std = ((G_d - G_d.mean(0)) / G_d.std(0, unbiased=False)).contiguous()

        def step_std():
            model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
            return stack5(model._scan_interaction_device(G_std))

        step_std()
        api.PROFILE.update(on=True, rot_ms=0.0, rot_flops=0.0, rot_flops_executed=0.0, rot_launches=0, int8_ms=0.0, int8_ops=0.0, int8_launches=0)
        ms_z, res_z = timed(step_std, max(2, a.steps // 2))
        api.PROFILE["on"] = False
        standardised = {"value": p_total / (ms_z / 1e3), "unit": UNIT, "ms_per_step": ms_z, "int8_contraction_launches": int(api.PROFILE["int8_launches"]),
                        "top_hits": torch.argsort(res_z[0])[:4].tolist(),
                        "note": "genotype columns standardised to mean 0, sd 1 (real-valued input): detected as affine images of integer dosages, "
                                "contracted on the int8 tensor cores and mapped back; not the headline value"}
        del G_std

    # ---- extension: donor-level genotype ingress; model kept across genes (N = 1 only) ----
    donor_level, shared_setup = None, None
    if extras and world == 1:
        Gdon_d = torch.from_numpy(Gd).to(dev)

        def step_donor():
            model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev, donor_level=True)
            return stack5(model._scan_interaction_device(Gdon_d, donor_index=donor_d))

        warm_up(step_donor, 2, at_most_extra=6)
        ms_d, res_d = timed(step_donor, max(2, a.steps // 2))
        pos = (res[0] > 0) & (res_d[0] > 0)
        donor_level = {"value": p_total / (ms_d / 1e3), "unit": UNIT, "ms_per_step": ms_d,
                       "max_abs_dlog10p_vs_expanded": float((torch.log10(res_d[0][pos]) - torch.log10(res[0][pos])).abs().max()),
                       "note": "same job with genotypes passed as a (donors x SNPs) matrix + donor index (keyword-only extension of the "
                               "reference API); the per-SNP contraction runs over donors instead of cells; not the headline value"}
        keep = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev)
        y_alt = [y_d + 0.01 * i * torch.sin(torch.arange(a.cells, device=dev, dtype=torch.float64)) for i in range(1, 3)]
        counter = {"i": 0}

        def step_shared():
            counter["i"] += 1
            keep.set_phenotype(y_alt[counter["i"] % 2])
            return stack5(keep._scan_interaction_device(G_d))

        warm_up(step_shared, 2, at_most_extra=6)
        ms_s, _ = timed(step_shared, max(2, a.steps // 2))
        shared_setup = {"value": p_total / (ms_s / 1e3), "ms_per_step": ms_s, "unit": UNIT,
                        "note": "model object kept across steps, CellRegMap.set_phenotype(y) per step (extension for scans of many genes "
                                "over the same cells); not the headline value"}
        del keep, Gdon_d

    # ---- weak scaling beside the strong-scaling headline: every rank scans its own p_total SNPs of the same gene ----
    weak = None
    if extras and strong:
        Gw_d = torch.from_numpy(donor_genotypes(a, rank, p_total) if rank else Gd).to(dev)[donor_d].contiguous()

        def step_weak():
            model = api._make_interaction_model(y_d, E_d, W_d, None, None, hK_d, device=dev, group=True)
            dist.all_gather_into_tensor(gathered_weak, stack5(model._scan_interaction_device(Gw_d)))
            return gathered_weak

        step_weak()
        ms_w, _ = timed(step_weak, max(2, a.steps // 2))
        weak = {"value": world * p_total / (ms_w / 1e3), "unit": UNIT, "ms_per_step": ms_w, "snps_per_gpu": p_total, "scaling": "weak"}
        del Gw_d

    # ---- e2e through the public API with plain numpy arrays (pageable host memory), what a user of the reference passes ----
    e2e = None
    if not a.no_e2e:
        del G_d
        torch.cuda.empty_cache()
        G_np = np.ascontiguousarray(Gd[gene["donor"]])         # (n, p_total) float64, pageable: the reference's asarray(G, float)
        y_np, W_np, E_np, hK_np = gene["y"], gene["W"], gene["E"], gene["hK"]

        def call_api(G):
            if a.entry == "run_association":
                return crm.run_association(y_np, W_np, E_np, G, hK=hK_np)[0]
            if a.entry == "estimate_betas":
                return crm.estimate_betas(y_np, W_np, E_np, G, hK=hK_np)[0]
            if world > 1 and strong:
                return crmd.run_interaction_sharded(y_np, E_np, G, W=W_np, hK=hK_np)[0]
            pv, info = crm.run_interaction(y_np, E_np, G, W=W_np, hK=hK_np)     # numpy results = D2H inside
            if world > 1:
                dist.all_gather_into_tensor(gathered_weak, torch.from_numpy(np.stack([pv, info["rho1"], info["e2"], info["g2"], info["eps2"]])).to(dev))
            return pv

        small = sum(x.size * 8 for x in (y_np, W_np, E_np, hK_np))
        cols = (hi - lo) if strong else p_total
        warm_up(lambda: call_api(G_np), 1, 6)
        ms_e, pv_e = timed(lambda: call_api(G_np), a.steps)
        e2e_units = units
        e2e = {"value": e2e_units / (ms_e / 1e3), "unit": UNIT, "ms_per_step": ms_e, "host_memory": "pageable numpy float64 (the reference's asarray(G, float))",
               "host_threads": int(lib.crm_host_threads()),
               # what crosses PCIe per step and rank: dosages as int8 after the host-side conversion (integer-valued input), the small arrays as float64
               "h2d_bytes_per_step": int(a.cells * cols + small), "host_bytes_read_per_step": int(a.cells * cols * 8 + small),
               "d2h_bytes_per_step": int(5 * cols * 8 if a.entry == "run_interaction" else cols * 8)}
        if a.entry == "run_interaction" and world == 1:
            assert np.array_equal(pv_e, res[0].cpu().numpy()), "host and device paths disagree"
        if extras:
            # the same call with page-locked float64 genotypes (DMA staging, 8 bytes per dosage over PCIe; N = 1 only) and with int8 dosages on
            # the host (1 byte per dosage to read, no conversion work for the host threads)
            arms = (("pinned_float64", lambda: torch.from_numpy(G_np).pin_memory(), "page-locked float64 host matrix: moved by DMA in column chunks"),
                    ("int8_host_genotypes", lambda: G_np.astype(np.int8), "dosages stored as int8 on the host (pageable)"))
            for key, G_alt, note in (arms if world == 1 else arms[1:]):
                try:
                    G_x = G_alt()
                except RuntimeError:
                    continue
                call_api(G_x)
                ms_x, pv_x = timed(lambda: call_api(G_x), max(2, a.steps // 2))
                assert np.array_equal(pv_x, pv_e) or world > 1, key
                e2e[key] = {"value": e2e_units / (ms_x / 1e3), "unit": UNIT, "ms_per_step": ms_x, "note": note}
                del G_x
        del G_np

    # ---- CPU baseline (rank 0, N = 1 only) ----
    os.sched_setaffinity(0, full_affinity)      # the CPU arm uses every host core again
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        os.environ.setdefault("TQDM_DISABLE", "1")
        ns = default_cpu_sample(a)
        r = cpu_reference_sample(a, gene, Gd, ns)
        v = ns / r["scan_s"][0]
        cpu = {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"first {ns} SNPs of the same gene, scan only ({r['scan_s'][0]:.1f} s); set-up ({r['setup_s']:.1f} s) excluded; code = {r['code']}"}

    if rank == 0:
        line = {"metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(a), "snps_total": units, "snps_per_gpu": (hi - lo) if strong else p_total,
                           "l2": "inputs (genotypes, basis) far larger than L2" if a.cells * a.snps * 8 > 4e8 else "small problem: inputs fit L2; every step rebuilds the model from its inputs",
                           "host_affinity": ("%d CPUs local to the GPU (NVML)" % len(local_cpus)) if local_cpus else "unchanged",
                           "step": "constructor set-up (shared between ranks) + scan of the rank's SNP block + all-gather of results"},
                "warmup_steps_run": warm_steps, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "fp64_route": fp64_route, "standardised_genotypes": standardised, "weak_scaling": weak,
                "cpu_baseline": cpu, "donor_level_ingress": donor_level, "shared_setup": shared_setup,
                "top_hits": top}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ctypes_double(fn):
    """value of a C-ABI probe `int fn(double* out, void* stream)`, None when it fails"""
    import ctypes
    v = ctypes.c_double(0.0)
    return float(v.value) if fn(ctypes.byref(v), None) == 0 else None


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
