"""Round 2: copy the evidence of the last GPU visits from gpurun_out/ (scratch) into profiles/ (tracked)."""
import collections, csv, json, os, shutil, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_profiles as MP

GO, PR, TAG = MP.GO, MP.PR, "r02"


def launch_table(src, dst, title):
    """Per-kernel table of the LAST complete step of an ncu launch list (steps are delimited by the forward pooling kernel)."""
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    items = []
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        items.append((r[ki], v, r[gi]))
    marks = [i for i, (n, _, _) in enumerate(items) if "pool_foldw_fwd" in n or "pool_fold_fwd" in n]
    # E > 8 runs the pooling in two entity passes: a step starts at every second mark then
    per = 2 if len(marks) >= 4 and marks[1] - marks[0] < 4 else 1
    marks = marks[::per]
    if len(marks) < 2:
        return
    step = items[marks[-2]:marks[-1]]
    short = lambda n: n.split("(")[0].replace("void ", "").replace("mvf::", "")
    agg = collections.OrderedDict()
    for n, v, g in step:
        a = agg.setdefault(short(n)[:72], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v for _, v, _ in step)
    with open(dst, "w") as f:
        f.write(f"{title}\nlaunches in one step: {len(step)}, sum of kernel durations {tot:.1f} us (serialised under ncu: the real step is "
                f"shorter, side-stream kernels overlap)\n\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v:9.1f} us {100 * v / tot:5.1f}% x{c:3d}  {k}\n")
        f.write("\n--- launch sequence (duration us, grid, kernel) ---\n")
        for n, v, g in step:
            f.write(f"{v:8.1f}  {g:>14s}  {short(n)[:72]}\n")


def to_bytes(x):
    num, _, unit = x.partition(" ")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(num.replace(",", "")) * mult


def traffic_json(vals, fname, kernel, workload, src):
    if vals and "dram__bytes_read.sum" in vals:
        tr = to_bytes(vals["dram__bytes_read.sum"]) + to_bytes(vals["dram__bytes_write.sum"])
        with open(os.path.join(PR, fname), "w") as f:
            json.dump({"kernel": kernel, "workload": workload, "dram_bytes_per_launch": tr,
                       "source": f"profiles/{src} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}, f, indent=1)


def main():
    cmd = "ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none; bench.py --workload {} --eager --steps 2 --warmup 3"
    for csvn, wl in (("launches.csv", "penn_cfg2"), ("launches4.csv", "finegym_cfg4"), ("launches5.csv", "long_cfg5")):
        launch_table(os.path.join(GO, csvn), os.path.join(PR, f"{TAG}_launches_{wl}.txt"), cmd.format(wl))
    S = "4 videos x 2 views, S = 16 x 240 = 3840, 8 heads x 32 channels (long_cfg5)"
    v = MP.ncu_summary("ncu_fa_fwd_kernel_long_cfg5.ncu-rep", os.path.join(PR, f"{TAG}_ncu_fa_fwd.txt"), f"fa_fwd_kernel<true>: tcgen05 flash attention forward, {S}")
    traffic_json(v, "attention_fwd_traffic.json", "fa_fwd_kernel<true> (long_cfg5)", "long_cfg5", f"{TAG}_ncu_fa_fwd.txt")
    MP.ncu_summary("ncu_fa_bwd_dq_kernel_long_cfg5.ncu-rep", os.path.join(PR, f"{TAG}_ncu_fa_bwd_dq.txt"), f"fa_bwd_dq_kernel<true>, {S}")
    MP.ncu_summary("ncu_fa_bwd_dkv_kernel_long_cfg5.ncu-rep", os.path.join(PR, f"{TAG}_ncu_fa_bwd_dkv.txt"), f"fa_bwd_dkv_kernel<true>, {S}")
    MP.ncu_summary("ncu_scl_pair_65536_20_128_0.ncu-rep", os.path.join(PR, f"{TAG}_ncu_scl_pair_scaled.txt"),
                   "scl_pair_mma_kernel<16,true,true,128,4>: 65 536 pairs x T 20 x D 128, all frames valid (scripts/scl_one.py)")
    MP.ncu_summary("ncu_scl_pair_8_80_256_1.ncu-rep", os.path.join(PR, f"{TAG}_ncu_scl_pair_cfg4_shape.txt"),
                   "scl_pair_mma_kernel<32,false,false,256,1> (thread-block clusters of 8): 8 pairs x T 80 x D 256, 20 % masked frames")
    v = MP.ncu_summary("ncu_pool_foldw_fwd_kernel_penn_cfg2.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_foldw_fwd.txt"),
                       "pool_foldw_fwd_kernel<9,true> at the bench shape (1280 frames x 196 tokens x 2304 ch, bf16, E=3)")
    traffic_json(v, "pool_fold_fwd_traffic.json", "pool_foldw_fwd_kernel<9,true> (cfg2: 1280 frames x 196 tokens x 2304 channels, bf16, E=3)",
                 "penn_cfg2", f"{TAG}_ncu_pool_foldw_fwd.txt")
    MP.ncu_summary("ncu_pool_foldw_bwd_kernel_penn_cfg2.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_foldw_bwd.txt"),
                   "pool_foldw_bwd_kernel<9,true> at the bench shape")
    for src, dst in (("scl_bench.txt", "scl_bench.txt"), ("gemm_bench.txt", "gemm_bench.txt"), ("gemm_dbg.txt", "gemm_dbg.txt"),
                     ("fa_dbg.txt", "fa_dbg.txt"), ("smoke.log", "smoke.log"), ("summary_tests.txt", "pytest_summary.txt")):
        if os.path.exists(os.path.join(GO, src)):
            shutil.copy(os.path.join(GO, src), os.path.join(PR, f"{TAG}_{dst}"))
    print("profiles/ updated:", sorted(n for n in os.listdir(PR) if n.startswith(TAG) or n.endswith("traffic.json")))


if __name__ == "__main__":
    main()
