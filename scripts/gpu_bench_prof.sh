set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:trace_step -s 4 -c 4 -o gpurun_out/prof_trace python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
