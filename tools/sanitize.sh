# compute-sanitizer over the hot path (the plane passes hand-roll mbarrier parity and per-warp TMA rings).
# usage (GPU box, repo root): bash tools/sanitize.sh <outdir>
O=${1:-gpurun_out/sanitize}
mkdir -p $O
T="tests/test_gpu_packed.py tests/test_gpu_batch.py tests/test_gpu_parity.py::test_summary_matches_oracle tests/test_gpu_sharded.py::test_peer_mailbox_exchange_two_ranks_in_one_process"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $T -m gpu -q -x -p no:cacheprovider > $O/$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$tool.log | tail -3 | tee -a $O/summary.txt
done
