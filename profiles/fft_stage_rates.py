"""Achieved FP32 throughput and bandwidth of the FFT-stage kernels from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv): per kernel the mean duration, the work one launch does
(from its grid), nominal flops (5 L log2 L per L-point complex FFT plus the per-bin stages) and algorithmic bytes, and the
resulting GFLOP/s and GB/s against the B200's 72 TFLOP/s FP32 (148 SMs x 128 lanes x 2 x 1.9 GHz) and the measured HBM peak.
  python profiles/fft_stage_rates.py gpurun_out/<run>/launches_c5.csv [more.csv ...] > profiles/r02_fft_stage_rates.json"""
import collections, csv, json, math, re, sys

PEAK_GBS = 6548.2
try:
    PEAK_GBS = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
FP32_TFLOPS = 148 * 128 * 2 * 1.9e9 / 1e12

def lines_per_cta(L):                       # LineTile<L>::C  (irb_spectral.cuh)
    G = 256 // (L // 8)
    return 4 if L >= 2048 else max(G, 8)

def work(name, grid):
    gx, gy, _ = grid
    m = re.search(r"k_fwd<(\d+)>", name)
    if m:                                   # rows of one M-point complex FFT + real split: 4B in (B = M), 8M out
        M = int(m.group(1)); rows = gx * (2048 // M)
        return rows, "rows", rows * (5 * M * math.log2(M) + 10 * M), rows * (4 * M + 8 * M)
    m = re.search(r"k_line_fft<(\d+), (\d)>", name)
    if m:                                   # lines of an L-point complex FFT (+ inter-pass twiddle): 16 L bytes per line
        L = int(m.group(1)); lines = gx * lines_per_cta(L) * gy
        return lines, "lines", lines * (5 * L * math.log2(L) + 6 * L), lines * 16 * L
    m = re.search(r"k_rowpair<(\d+)>", name)
    if m:                                   # rows: forward + inverse L-point FFT, split, multiply, merge, twiddle: 16 L bytes per row (+ the shared spectrum B)
        L = int(m.group(1)); NP = max(1, (256 // (L // 8)) // 2); rows = 2 * gx * NP * gy
        return rows, "rows", rows * (2 * 5 * L * math.log2(L) + 40 * L), rows * 16 * L
    return None

acc = collections.OrderedDict()
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]
    ik, ig, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    for r in rows[1:]:
        name = r[ik].replace("irb::", "").replace("void ", "")
        grid = tuple(int(v) for v in re.findall(r"\d+", r[ig]))
        w = work(name, grid)
        if not w:
            continue
        key = (name.split("(")[0], grid)
        a = acc.setdefault(key, {"ns": [], "w": w})
        a["ns"].append(float(r[iv].replace(",", "")))
out = []
for (name, grid), a in acc.items():
    n, unit, flops, byts = a["w"]
    ns = sorted(a["ns"])
    t = ns[len(ns) // 2] * 1e-9
    out.append({"kernel": name, "grid": list(grid), "launches": len(ns), "median_us": round(t * 1e6, 2), unit: n, "nominal_gflop_per_launch": round(flops / 1e9, 4),
                "algorithmic_mb_per_launch": round(byts / 1e6, 3), "achieved_tflops_fp32": round(flops / t / 1e12, 2), "frac_of_fp32_peak": round(flops / t / 1e12 / FP32_TFLOPS, 3),
                "achieved_gbs": round(byts / t / 1e9, 1), "frac_of_measured_hbm_peak": round(byts / t / 1e9 / PEAK_GBS, 3)})
json.dump({"note": "per-launch times under ncu are cold-cache and serialised (an upper bound of the in-pipeline time); fp32 peak %.1f TFLOP/s nominal, HBM peak %.1f GB/s measured" % (FP32_TFLOPS, PEAK_GBS),
           "kernels": out}, sys.stdout, indent=1)
