"""Copy the evidence of the last GPU visits from gpurun_out/ (scratch, git-ignored) into profiles/ (tracked):
ncu launch lists -> per-kernel tables, ncu --set full captures -> key metrics, micro-benchmarks and bench lines."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def launch_table(src, dst, title):
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    items = []
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        items.append((r[ki], v, r[gi]))
    names = [n for n, _, _ in items]
    marks = [i for i, n in enumerate(names) if "posenc_table" in n]
    if len(marks) < 2:
        return
    def start(i):
        while i > 0 and "pack_kernel" not in names[i]:
            i -= 1
        return i
    s, e = start(marks[-2]), start(marks[-1])
    step = items[s:e]
    short = lambda n: n.split("(")[0].replace("void ", "").replace("mvf::", "")
    agg = collections.OrderedDict()
    for n, v, g in step:
        a = agg.setdefault(short(n)[:70], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v for _, v, _ in step)
    with open(dst, "w") as f:
        f.write(f"{title}\nlaunches in one step: {len(step)}, sum of kernel durations {tot:.1f} us (serialised under ncu)\n\n")
        for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v:9.1f} us {100 * v / tot:5.1f}% x{c:3d}  {k}\n")
        f.write("\n--- launch sequence (duration us, grid, kernel) ---\n")
        for n, v, g in step:
            f.write(f"{v:8.1f}  {g:>14s}  {short(n)[:70]}\n")


def ncu_summary(rep, dst, title):
    src = os.path.join(GO, rep)
    if not os.path.exists(src):
        return None
    det = subprocess.run(["ncu", "-i", src, "--page", "details"], capture_output=True, text=True).stdout
    keep = ("Duration", "Elapsed Cycles", "SM Frequency", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput",
            "Compute (SM) Throughput", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Dynamic Shared Memory",
            "Grid Size", "Block Size", "Achieved Occupancy", "L2 Hit Rate", "No Eligible", "Warp Cycles Per Issued", "Mem Busy",
            "Max Bandwidth", "Waves Per SM")
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    vals = {}
    if len(rr) >= 3:
        for n, u, v in zip(rr[0], rr[1], rr[2]):
            vals[n] = f"{v} {u}".strip()
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.sum",
            "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    with open(dst, "w") as f:
        f.write(title + "\n(ncu --set full --clock-control none --import-source on, one launch; numbers under the profiler are NOT bench values)\n\n")
        for ln in det.splitlines():
            if ln.strip().startswith("void ") or ln.strip().startswith("mvf::") or "Context" in ln and "Stream" in ln:
                f.write(ln.rstrip() + "\n")
            elif any(ln.strip().startswith(k) for k in keep):
                f.write(ln.rstrip() + "\n")
        f.write("\nraw metrics:\n")
        for w in want:
            if w in vals:
                f.write(f"  {w} = {vals[w]}\n")
    return vals


def unit_bytes(v, unit_row):
    return v


def main():
    os.makedirs(PR, exist_ok=True)
    launch_table(os.path.join(GO, "launches.csv"), os.path.join(PR, f"{TAG}_launches_step_cold.txt"),
                 "ncu --metrics gpu__time_duration.sum --clock-control none; bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense (cfg2, bf16, folded pooling)")
    launch_table(os.path.join(GO, "launches_warm.csv"), os.path.join(PR, f"{TAG}_launches_step_warm.txt"),
                 "same with --cache-control none (L2 state carried between kernels, as in the real step)")
    ncu_summary("prof_fold_fwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_fold_fwd.txt"), "pool_foldm_fwd_kernel<12> (first generation, MVF_FOLD_WS=0) at the bench shape (1280 frames x 196 tokens x 2304 ch, bf16, E=3)")
    ncu_summary("prof_fold_bwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_fold_bwd.txt"), "pool_foldm_bwd_kernel<12> (first generation) at the bench shape")
    v = ncu_summary("prof_foldw_fwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_foldw_fwd.txt"), "pool_foldw_fwd_kernel<9,true> (warp-specialised, default) at the bench shape (1280 frames x 196 tokens x 2304 ch, bf16, E=3)")
    ncu_summary("prof_foldw_bwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_pool_foldw_bwd.txt"), "pool_foldw_bwd_kernel<9,true> at the bench shape")
    ncu_summary("prof_attn_tc_bwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_attn_tc_bwd.txt"), "attn_tc_bwd_kernel (64 views x 8 heads, S = 60, d_k = 32)")
    ncu_summary("prof_gemm_split3.ncu-rep", os.path.join(PR, f"{TAG}_ncu_gemm_bf16x3_3840x512x512.txt"), "gemm_tc_kernel<256,4,4,true> (bf16x3, 3840 x 512 x 512)")
    ncu_summary("prof_kv_fwd.ncu-rep", os.path.join(PR, f"{TAG}_ncu_kv_proj_fwd_dense.txt"), "gemm_tc_kernel<256,4,2,false> dense K|V projection forward (250880 x 768 x 2304, bf16)")
    for name in ("fold_bench.txt", "scl_bench.txt", "bench_2gpu_graph.json", "bench_8gpu.json", "bench_reference.json", "host_cost.txt", "gemm_dbg.txt", "gemm_bench.txt", "diag_cfg1.txt", "bench.json",
                 "bench_quick.json", "bench_2gpu.json", "pytest_summary.txt", "smoke.log"):
        src = os.path.join(GO, name)
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PR, f"{TAG}_{name}"))
    if v and "dram__bytes_read.sum" in v:
        def to_bytes(x):
            num, _, unit = x.partition(" ")
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
            return float(num.replace(",", "")) * mult
        tr = to_bytes(v["dram__bytes_read.sum"]) + to_bytes(v["dram__bytes_write.sum"])
        with open(os.path.join(PR, "pool_fold_fwd_traffic.json"), "w") as f:
            json.dump({"kernel": "pool_foldw_fwd_kernel<9,true> (cfg2: 1280 frames x 196 tokens x 2304 channels, bf16, E=3)",
                       "dram_bytes_per_launch": tr,
                       "source": f"profiles/{TAG}_ncu_pool_foldw_fwd.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}, f, indent=1)
    print("profiles/ updated:", sorted(os.listdir(PR)))


if __name__ == "__main__":
    main()
