"""Prints the BASELINE.md table rows from the bench lines kept under profiles/ (r02_bench_*.json)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    try:
        return json.load(open(os.path.join(HERE, name)))
    except (OSError, ValueError):
        return None


n1 = load("r02_bench_n1.json")
rows = []
for cfg, name in ((1, "r02_bench_cfg1.json"), (2, "r02_bench_cfg2.json"), (3, "r02_bench_n1.json"), (4, "r02_bench_cfg4.json"), (5, "r02_bench_cfg5.json")):
    d = load(name)
    if d is None:
        continue
    cpu = d.get("cpu_baseline") or {}
    e2e = d.get("e2e") or {}
    roof = d.get("roofline") or {}
    multi = {}
    if cfg == 3:
        for n in (2, 4, 8):
            m = load(f"r02_bench_n{n}.json")
            if m:
                multi[n] = m
    cells = [str(cfg), d["metric"].split("(")[1].rstrip(")"),
             f"{cpu.get('value', float('nan')):.3g} ({cpu.get('cores', '?')} cores)",
             f"{d['value']:,.0f}", f"{e2e.get('value', float('nan')):,.0f}"]
    for n in (2, 4, 8):
        m = multi.get(n)
        cells.append(f"{m['value']:,.0f} / {m['e2e']['value']:,.0f} / {m['weak_scaling']['value']:,.0f}" if m else "–")
    cells.append(f"{roof['frac']:.2f} ({roof['unit']})" if roof.get("frac") else "–")
    rows.append("| " + " | ".join(cells) + " |")
print("| config | entry point, size | reference CPU tests/s | GPU ×1 device-resident | ×1 e2e (pageable numpy) | ×2 strong / e2e / weak | ×4 | ×8 | roofline fraction ×1 |")
print("|---|---|---|---|---|---|---|---|---|")
print("\n".join(rows))
if n1:
    print()
    for key in ("fp64_route", "standardised_genotypes", "donor_level_ingress", "shared_setup"):
        v = n1.get(key)
        if v:
            print(f"* `{key}`: {v['value']:,.0f} tests/s ({v['ms_per_step']:.1f} ms per step)")
    e = n1["e2e"]
    print(f"* e2e sub-arms: pinned float64 {e['pinned_float64']['value']:,.0f}, int8 host genotypes {e['int8_host_genotypes']['value']:,.0f} tests/s")
