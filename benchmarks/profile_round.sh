#!/bin/bash
# Everything profiles/<round>/ is built from, in one GPU-box call:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash benchmarks/profile_round.sh'
# Numbers printed by a run under ncu are never bench values (those come from bench.py / benchmarks/*.py with CUDA events).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py > $O/bench.json 2> $O/bench.err
python benchmarks/kernels.py > $O/kernels.jsonl 2> $O/kernels.err
python benchmarks/host_path.py > $O/host_path.jsonl 2> $O/host_path.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --preheat 0 --no-cpu-baseline > $O/ncu_bench.log 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:pose_solve -s 2 -c 1 -f -o $O/prof_solve \
    python benchmarks/prof_solve.py 1024 64 > $O/ncu_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pull_gated -s 8 -c 1 -f -o $O/prof_pull \
    python benchmarks/host_path.py > $O/ncu_pull.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:correspond -s 2 -c 1 -f -o $O/prof_s1 \
    python benchmarks/kernels.py --only s1 > $O/ncu_s1.log 2>&1
for tool in memcheck racecheck synccheck; do
    echo "== $tool" >> $O/sanitizer.txt
    compute-sanitizer --tool $tool python benchmarks/sanitize.py 2>&1 | tail -3 >> $O/sanitizer.txt
done
tail -c 400 $O/bench.json
