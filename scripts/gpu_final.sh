# Round-end evidence run: tests, both bench arms, launch list, ncu --set full of the three dominant kernels.
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"intersect_wave|interact_wave" -s 24 -c 8 -f -o gpurun_out/${TAG}_trace python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-detector > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"pd_field_fast|compact_scatter" -c 3 -f -o gpurun_out/${TAG}_pd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --c3-side 16 > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu pd rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
