"""python profiles/launch_shares.py <ncu launch list csv>: kernel time per kernel name over the last step of the run (a step starts with
the QR of the context matrix in the Python layer) -- the shares that bench.py's roofline.share_of_step must agree with."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
    elif hdr and r[0].isdigit():
        data.append(dict(zip(hdr, r)))
starts = [i for i, d in enumerate(data) if "geqr2" in d["Kernel Name"]]
seg = data[starts[-1]:] if starts else data
seg = [d for d in seg if "crm_dmma_rate_kernel" not in d["Kernel Name"]]      # the FP64 peak probe bench.py runs after the timed steps
tot, cnt = collections.Counter(), collections.Counter()
for d in seg:
    name = re.sub(r"\(.*", "", d["Kernel Name"])[:70]
    v = float(d["Metric Value"].replace(",", ""))
    ms = v / 1e6 if d["Metric Unit"].startswith("n") else v / 1e3 if d["Metric Unit"].startswith("u") else v
    tot[name] += ms
    cnt[name] += 1
s = sum(tot.values())
ours = sum(c for k, c in cnt.items() if "crm::" in k)
print(f"last step of {sys.argv[1]}: {len(seg)} launches ({ours} of libcrm_b200 kernels), kernel time {s:.1f} ms (cold-cache, serialised under ncu)")
for k, v in tot.most_common(24):
    print(f"{v:8.2f} ms {100 * v / s:5.1f}%  x{cnt[k]:<4d} {k}")
