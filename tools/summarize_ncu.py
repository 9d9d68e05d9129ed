"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...` launch list per kernel family (run here, no GPU needed)."""
import collections, csv, re, sys

def load(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    idx = {n: i for i, n in enumerate(hdr)}
    data = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or not r[idx["ID"]].isdigit():
            continue
        d = data.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]]})
        try:
            d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    return data

def family(name):
    m = re.search(r"(gemm_tc_kernel<[^>]*>|[a-z_0-9]+_kernel)", name)
    if m:
        return m.group(1)
    return name.split("(")[0][-40:]

def main(path):
    data = load(path)
    fam = collections.OrderedDict()
    tot = 0.0
    for d in data.values():
        f = fam.setdefault(family(d["name"]), {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "tensor_w": 0.0})
        ns = d.get("gpu__time_duration.sum", 0.0)
        f["n"] += 1
        f["ns"] += ns
        f["rd"] += d.get("dram__bytes_read.sum", 0.0)
        f["wr"] += d.get("dram__bytes_write.sum", 0.0)
        f["tensor_w"] += ns * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        tot += ns
    print(f"{path}: {len(data)} launches, {tot / 1e6:.3f} ms (serialised, cold cache: compare shares)")
    print(f"{'family':44s} {'n':>4s} {'ms':>8s} {'share':>6s} {'GB/s':>8s} {'tensor%':>8s}")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
        gbs = (f["rd"] + f["wr"]) / max(f["ns"], 1)
        print(f"{k:44s} {f['n']:4d} {f['ns'] / 1e6:8.3f} {f['ns'] / tot:6.1%} {gbs:8.1f} {f['tensor_w'] / max(f['ns'], 1):8.1f}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
