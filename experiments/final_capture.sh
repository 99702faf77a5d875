set -x
mkdir -p gpurun_out/final
timeout 300 python benchmarks/bench_stream.py > gpurun_out/final/bench_stream.log 2>&1
tail -3 gpurun_out/final/bench_stream.log
timeout 300 python benchmarks/bench_em.py 6 > gpurun_out/final/bench_em.json 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > gpurun_out/final/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gmm_score_tc2 -s 2 -c 1 -o gpurun_out/final/score_tc2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mfcc_frames_warp -s 2 -c 1 -o gpurun_out/final/mfcc_warp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > /dev/null 2>&1
ls -la gpurun_out/final
