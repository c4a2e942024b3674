TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v Warn | tail -20 > gpurun_out/mgpu_check_n8_r2.txt; tail -3 gpurun_out/mgpu_check_n8_r2.txt
$TR --master-port 29512 tools/shard_profile.py 2>&1 | grep -v Warn | tail -11 | tee gpurun_out/shard_profile_n8_r2.txt
$TR --master-port 29513 tools/pcie_bench.py 2>&1 | grep "^{" | tee gpurun_out/pcie_n8_r2.json
python tools/pcie_bench.py 2>&1 | grep "^{" | tee gpurun_out/pcie_n1_r2.json
$TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_s20_r2.json 2> gpurun_out/bench_n8_s20_r2.err
$TR --master-port 29515 bench.py --gpus 8 --steps 19 --warmup 3 --width 3840 --height 2160 --frames 2500 --e2e-steps 8 > gpurun_out/bench_4k_n8_r2.json 2> gpurun_out/bench_4k_n8_r2.err
python - <<PY
import json
for f in ("gpurun_out/bench_n8_s20_r2.json","gpurun_out/bench_4k_n8_r2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["resident_frames_per_gpu"])
    except Exception as e: print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
