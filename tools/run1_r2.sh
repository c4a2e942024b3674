# round-2 single-GPU evidence run:   gpurun --timeout 1500 -- 'bash tools/run1_r2.sh'
# every profiled command first exits 0 without ncu; numbers printed under ncu are never bench values
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gputest_r2_final.log
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r2_n1_steps20.json 2>/dev/null
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_reference.json 2>/dev/null
python tools/kbench.py --batch 128 --iters 30 > gpurun_out/kbench_r2.txt 2>&1
python tools/kbench.py --w 640 --h 480 --batch 256 --iters 30 > gpurun_out/kbench_vga_r2.txt 2>&1
python tools/label_bench.py --batch 128 2>&1 | grep mask > gpurun_out/label_bench_r2.txt
python tools/label_bench.py --w 640 --h 480 --batch 256 2>&1 | grep mask >> gpurun_out/label_bench_r2.txt
python tools/blur_bench.py --batch 128 --sigmas 2 > gpurun_out/blur_bench_r2.txt 2>&1
python tools/blur_bench.py --batch 64 >> gpurun_out/blur_bench_r2.txt 2>&1
python tools/configs_bench.py > gpurun_out/configs_r2.txt 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/launches_r2.log 2>&1
PROF_BATCH=128 python tools/prof_once.py > /dev/null 2>&1 && \
PROF_BATCH=128 ncu --set full --clock-control none -s 19 -c 26 -f -o gpurun_out/prof_all_r2c python tools/prof_once.py > gpurun_out/prof_all_r2c.log 2>&1
ls -la gpurun_out/prof_all_r2c.ncu-rep
tail -3 gpurun_out/gputest_r2_final.log
