#!/bin/bash
# One GPU-box visit of round 2.  Usage: scripts/gpu_r2.sh [tests] [smoke] [bench] [bench4] [bench5] [ref] [launches] ...
# Everything is logged under gpurun_out/ (merged back by gpurun).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
S=gpurun_out/summary.txt
for stage in "$@"; do
  echo "=== $stage === $(date +%T)" | tee -a $S
  case $stage in
    tests)
      for f in test_gpu_gemm test_gpu_fold test_gpu_blocks test_gpu_scl test_gpu_model test_gpu_dropin test_gpu_graph test_gpu_optim test_gpu_multi test_gpu_fullsize; do
        [ -f tests/$f.py ] || continue
        timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=1200 -p no:cacheprovider > gpurun_out/$f.log 2>&1
        echo "$f exit $? :: $(tail -n 1 gpurun_out/$f.log)" | tee -a $S
      done ;;
    tests_extra)
      for f in $(ls tests/test_gpu_*.py | sed 's#tests/##; s#\.py##'); do
        case $f in test_gpu_gemm|test_gpu_fold|test_gpu_blocks|test_gpu_scl|test_gpu_model|test_gpu_dropin|test_gpu_graph|test_gpu_optim|test_gpu_multi|test_gpu_fullsize) continue;; esac
        timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=1200 -p no:cacheprovider > gpurun_out/$f.log 2>&1
        echo "$f exit $? :: $(tail -n 1 gpurun_out/$f.log)" | tee -a $S
      done ;;
    file:*)
      f=${stage#file:}
      timeout 1500 python -m pytest tests/$f.py -m gpu -q -s --timeout=1200 -p no:cacheprovider > gpurun_out/$f.log 2>&1
      echo "$f exit $? :: $(tail -n 1 gpurun_out/$f.log)" | tee -a $S ;;
    smoke)
      timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a $S
      tail -n 4 gpurun_out/smoke.log | tee -a $S ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "exit $?" | tee -a $S ;;
    bench4)
      timeout 900 python bench.py --workload finegym_cfg4 --steps 20 --warmup 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "exit $?" | tee -a $S ;;
    bench5)
      timeout 1200 python bench.py --workload long_cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "exit $?" | tee -a $S ;;
    quick)
      timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-dense --no-refgpu --no-parity > gpurun_out/quick_cfg2.json 2> gpurun_out/quick_cfg2.err; echo "exit $?" | tee -a $S ;;
    quick4)
      timeout 600 python bench.py --workload finegym_cfg4 --steps 20 --warmup 5 --no-e2e --no-cpu --no-dense --no-refgpu --no-parity > gpurun_out/quick_cfg4.json 2> gpurun_out/quick_cfg4.err; echo "exit $?" | tee -a $S ;;
    quick5)
      timeout 600 python bench.py --workload long_cfg5 --steps 10 --warmup 3 --no-e2e --no-cpu --no-dense --no-refgpu --no-parity > gpurun_out/quick_cfg5.json 2> gpurun_out/quick_cfg5.err; echo "exit $?" | tee -a $S ;;
    diag)
      DIAG_ALL=1 timeout 900 python scripts/diag_fullsize.py 32 20 3 512 > gpurun_out/diag_fullsize.txt 2>&1
      MVF_SIMT_CHUNK=0 timeout 900 python scripts/diag_fullsize.py 32 20 3 512 >> gpurun_out/diag_fullsize.txt 2>&1
      echo "exit $?" | tee -a $S; grep "grad rel" gpurun_out/diag_fullsize.txt | tee -a $S ;;
    ncu:*)
      # ncu:<kernel regex>:<workload>  -> one full capture of the matching kernel
      spec=${stage#ncu:}; kre=${spec%%:*}; wl=${spec#*:}
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 2 -f -o gpurun_out/ncu_${kre}_$wl \
        python bench.py --workload $wl --eager --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense --no-refgpu --no-parity > gpurun_out/ncu_${kre}_$wl.log 2>&1
      echo "exit $?" | tee -a $S ;;
    dist|dist4|dist5)
      # N = every GPU of the box: the driver's launch line; variants: default, nooverlap, nccl, ctas8 (DIST_VARIANTS)
      n=$(nvidia-smi -L | wc -l); wl=penn_cfg2; [ $stage = dist4 ] && wl=finegym_cfg4; [ $stage = dist5 ] && wl=long_cfg5
      for v in ${DIST_VARIANTS:-default nooverlap nccl}; do
        tag=${wl}_n${n}_$v
        ar=1; [ "$v" = nccl ] && ar=0
        xo=""; [ "$v" = nooverlap ] && xo="--no-overlap"
        ctas=0; [ "$v" = ctas8 ] && ctas=8
        MVF_PEER_AR_CTAS=$ctas MVF_PEER_AR=$ar timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py \
          --gpus $n --workload $wl --steps 50 --warmup 5 --no-cpu --no-refgpu --no-dense $xo ${DIST_EXTRA:-} > gpurun_out/dist_$tag.json 2> gpurun_out/dist_$tag.err
        echo "$tag exit $? :: $(tail -c 400 gpurun_out/dist_$tag.json | tr '\n' ' ' | grep -o '"parity".*' | head -c 300)" | tee -a $S
        python - gpurun_out/dist_$tag.json <<'PY' | tee -a $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   n_gpus", d["n_gpus"], "value %.1f ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", (d.get("e2e") or {}).get("value"), "cross_rank", d.get("cross_rank"))
except Exception as e:
    print("   unreadable:", e)
PY
      done ;;
    scl)
      timeout 300 python scripts/scl_bench.py > gpurun_out/scl_bench.txt 2>&1; echo "exit $?" | tee -a $S
      python - <<'PY' | tee -a $S
import json
for l in open("gpurun_out/scl_bench.txt"):
    try:
        d = json.loads(l)
        print("   pairs %6d T %3d D %3d masked %-5s  %8.4f ms  %7.1f GB/s  frac %.3f" % (d["pairs"], d["T"], d["D"], d["masked_frames"], d["ms"], d["GBps"], d["frac_of_measured_hbm"]))
    except Exception:
        pass
PY
      ;;
    sclncu:*)
      # sclncu:<Bv>:<T>:<D>:<masked>  launch list + full capture of the pair kernel for one SCL shape
      spec=${stage#sclncu:}; IFS=: read bv tt dd mm <<< "$spec"
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/scl_launches_${bv}_${tt}_${dd}_${mm}.csv python scripts/scl_one.py $bv $tt $dd $mm > /dev/null 2>&1
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:scl_pair_mma -s 2 -c 1 -f -o gpurun_out/ncu_scl_pair_${bv}_${tt}_${dd}_${mm} python scripts/scl_one.py $bv $tt $dd $mm > gpurun_out/ncu_scl_pair_${bv}_${tt}_${dd}_${mm}.log 2>&1
      echo "exit $?" | tee -a $S
      python - gpurun_out/scl_launches_${bv}_${tt}_${dd}_${mm}.csv <<'PY' | tee -a $S
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
for r in rows[1:][-8:]:
    print("   %-70s %s" % (r[ki][:70], r[vi]))
PY
      ;;
    peer)
      n=$(nvidia-smi -L | wc -l)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 scripts/peer_bench.py > gpurun_out/peer_bench_n$n.json 2> gpurun_out/peer_bench_n$n.err
      echo "exit $? :: $(tail -n 1 gpurun_out/peer_bench_n$n.json)" | tee -a $S ;;
    strong)
      # BASELINE configs[2]: 256 Penn videos in total, split over the box's GPUs (strong scaling)
      n=$(nvidia-smi -L | wc -l)
      if [ $n -gt 1 ]; then
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py \
          --gpus $n --global-videos 256 --steps 20 --warmup 5 --no-cpu --no-refgpu --no-dense --no-e2e > gpurun_out/strong_n$n.json 2> gpurun_out/strong_n$n.err
      else
        timeout 900 python bench.py --global-videos 256 --steps 20 --warmup 5 --no-cpu --no-refgpu --no-dense --no-e2e > gpurun_out/strong_n$n.json 2> gpurun_out/strong_n$n.err
      fi
      echo "exit $?" | tee -a $S
      python - gpurun_out/strong_n$n.json <<'PY' | tee -a $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   n_gpus", d["n_gpus"], "videos/GPU", d["config"]["videos_per_gpu"], "value %.1f ms %.3f" % (d["value"], d["ms_per_step"]), d["scaling"], "parity ok", (d.get("parity") or {}).get("ok"))
except Exception as e:
    print("   unreadable:", e)
PY
      ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "exit $?" | tee -a $S ;;
    launches|launches4|launches5)
      wl=penn_cfg2; [ $stage = launches4 ] && wl=finegym_cfg4; [ $stage = launches5 ] && wl=long_cfg5
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/$stage.csv \
        python bench.py --workload $wl --eager --steps 2 --warmup 3 --no-e2e --no-cpu --no-dense --no-refgpu --no-parity > gpurun_out/$stage.log 2>&1
      echo "exit $?" | tee -a $S ;;
    *) echo "unknown stage $stage" | tee -a $S ;;
  esac
done
for f in gpurun_out/bench_cfg*.json gpurun_out/quick_cfg*.json; do [ -f $f ] && python - "$f" <<'PY' | tee -a $S
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "value %.1f ms %.3f launches %s | roofline %s frac %.3f | parity %s | e2e %s | refgpu %s" % (
        d["value"], d["ms_per_step"], d.get("gpu_launches"), (r.get("kernel") or "")[:40], r.get("frac") or 0,
        {k: (round(v, 6) if isinstance(v, float) else v) for k, v in (d.get("parity") or {}).items() if k in ("loss_rel", "emb_rel", "grad_rel", "ok")},
        (d.get("e2e") or {}).get("value"), {k: round(v["value"], 1) for k, v in (d.get("reference_gpu") or {}).items() if isinstance(v, dict) and "value" in v}))
    print("   share", {k: round(v, 3) for k, v in (r.get("share_of_step") or {}).items() if v})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
exit 0
