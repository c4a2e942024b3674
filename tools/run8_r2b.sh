# round-2 multi-GPU evidence run after the kernel / egress work (8 GPUs of one box):   gpurun --gpus 8 -- 'bash tools/run8_r2b.sh'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -v Warn | tail -20 > gpurun_out/mgpu_check_n8_r2.txt; tail -1 gpurun_out/mgpu_check_n8_r2.txt
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_s20_r2.json 2> gpurun_out/bench_n8_s20_r2.err
$TR --nproc-per-node 4 --master-port 29516 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4_s20_r2.json 2> gpurun_out/bench_n4_s20_r2.err
$TR --nproc-per-node 2 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_s20_r2.json 2> gpurun_out/bench_n2_s20_r2.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n1_s20_r2.json 2>/dev/null
$TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 19 --warmup 3 --width 3840 --height 2160 --frames 2500 --e2e-steps 8 > gpurun_out/bench_4k_n8_r2.json 2> gpurun_out/bench_4k_n8_r2.err
python - <<PY
import json
for f in ("gpurun_out/bench_n1_s20_r2.json","gpurun_out/bench_n2_s20_r2.json","gpurun_out/bench_n4_s20_r2.json","gpurun_out/bench_n8_s20_r2.json","gpurun_out/bench_4k_n8_r2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["runs"], d["e2e_region_table"]["runs"], d["e2e_dense_copy"]["runs"])
    except Exception as e: print(f, "ERR", e)
PY
