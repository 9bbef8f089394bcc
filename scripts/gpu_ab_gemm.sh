#!/bin/bash
# A/B of the GEMM tile height: parity of both variants, then the bench with MH pinned to 1 and chosen by shape.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mask_head.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_mask_head_mh.log 2>&1
echo "test_gpu_mask_head exit=$?" | tee gpurun_out/summary_ab.txt
tail -n 15 gpurun_out/test_mask_head_mh.log
timeout 200 python -m pytest tests/test_gpu_net.py tests/test_gpu_att.py -x -q -m gpu -p no:cacheprovider > gpurun_out/test_net_mh.log 2>&1
echo "test_gpu_net/att exit=$?" | tee -a gpurun_out/summary_ab.txt
tail -n 5 gpurun_out/test_net_mh.log
L2S_GEMM_MH=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mh1.json 2> gpurun_out/bench_mh1.err
echo "bench mh1 exit=$?" | tee -a gpurun_out/summary_ab.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mhauto.json 2> gpurun_out/bench_mhauto.err
echo "bench auto exit=$?" | tee -a gpurun_out/summary_ab.txt
python - <<'PY'
import json
for t in ("mh1", "mhauto"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % t))
        print(t, "ms/step %.3f" % d["ms_per_step"], "value %.1f" % d["value"], "clocks", d["clocks"])
        for c in d["components"]:
            if "mask" in c["kernel"] or "gemm" in c["kernel"]:
                print("   ", c["kernel"][:50], c["ms"], c["frac"])
    except Exception as e:
        print(t, "failed", e)
PY
