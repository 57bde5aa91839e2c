# bench lines of the other BASELINE configs (device-resident legs only, no CPU baseline: chunkyseal costs ~20 s per frame on the host)
mkdir -p gpurun_out
( time timeout 300 python bench.py --card chunkyseal --batch 16 --size 512 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out gpurun_out/prof_chunky.json > gpurun_out/bench_chunky.json ) 2> gpurun_out/bench_chunky.err; tail -c 400 gpurun_out/bench_chunky.json; tail -5 gpurun_out/bench_chunky.err
( time timeout 90 python bench.py --card pixelseal --batch 32 --size 768 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_pixelseal768.json ) 2> gpurun_out/bench_pixelseal768.err; tail -c 300 gpurun_out/bench_pixelseal768.json
( time timeout 90 python bench.py --card videoseal_1.0 --batch 64 --size 768 --video --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_v1_video768.json ) 2> gpurun_out/bench_v1_video768.err; tail -c 300 gpurun_out/bench_v1_video768.json
