set -x
mkdir -p gpurun_out/r1b
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r1b/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1b/pytest.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r1b/bench.json 2> gpurun_out/r1b/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1b/bench_ref.json 2> gpurun_out/r1b/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b/launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/r1b/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fm_k_plane_pass -s 8 -c 2 -o gpurun_out/r1b/prof_plane_v5 -f python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/r1b/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/r1b/pytest.log; cat gpurun_out/r1b/bench.json
