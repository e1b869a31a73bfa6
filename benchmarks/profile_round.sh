#!/bin/bash
# Everything profiles/<round>/ is built from, in one GPU-box call:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash benchmarks/profile_round.sh'
# Numbers printed by a run under ncu are never bench values (those come from bench.py / benchmarks/*.py with CUDA events).
set -u
O=gpurun_out/r2
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --steps 20 --warmup 3 > $O/bench_1gpu.json 2> $O/bench.err
python benchmarks/pipeline.py --configs ycbv,ycbv4k,ycbv2k,lmo --chunks 0 > $O/pipeline.jsonl 2> $O/pipeline.err
python benchmarks/pipeline.py --configs ycbv,lmo --chunks 0 --streams 2 >> $O/pipeline.jsonl 2>> $O/pipeline.err
python benchmarks/kernels.py > $O/kernels.jsonl 2> $O/kernels.err
python benchmarks/fps_small.py > $O/fps_small.jsonl 2> $O/fps_small.err
python benchmarks/fps_ppt_probe.py > $O/fps_ppt_probe.txt 2> $O/fps_ppt_probe.err
python benchmarks/host_path.py > $O/host_path.jsonl 2> $O/host_path.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 --preheat 0 --no-cpu-baseline > $O/ncu_bench.log 2>&1
# full captures of the kernels of the headline workload (YCB-V 8192) and of S1 on the same maps
for k in front_kernel score_kernel refit_kernel correspond_kernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/prof_$k \
        python benchmarks/prof_pipeline.py ycbv split 0 > $O/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:pose_solve_kernel -s 2 -c 1 -f -o $O/prof_pose_solve_kernel \
    python benchmarks/prof_pipeline.py ycbv fused 0 > $O/ncu_pose_solve_kernel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fps_cluster -s 2 -c 1 -f -o $O/prof_fps_cluster \
    python benchmarks/fps_small.py > $O/ncu_fps_cluster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coor_feat -s 3 -c 1 -f -o $O/prof_coor_feat \
    python benchmarks/kernels.py --only misc --no-cpu > $O/ncu_coor_feat.log 2>&1
for tool in memcheck racecheck synccheck; do
    echo "== $tool" >> $O/sanitizer.txt
    compute-sanitizer --tool $tool python benchmarks/sanitize.py 2>&1 | tail -3 >> $O/sanitizer.txt
done
tail -c 300 $O/bench_1gpu.json
