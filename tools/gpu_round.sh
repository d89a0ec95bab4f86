# One GPU round: tests, bench (both arms), ncu launch list and one full capture of the top kernel.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag>
set -x
T=${1:-r}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
if [ -z "$NO_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > $O/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fm_k_plane_pass -s 4 -c 1 -o $O/prof_plane -f python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
cat $O/bench.json
