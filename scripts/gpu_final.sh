# Round-end evidence run: tests, smoke, both bench arms, full-size configs, launch list, ncu --set full of the dominant kernels.
# Every step under `timeout`.  The .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): their summaries are written here.
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python scripts/bench_configs.py --c3-side 32 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --c3-side 16"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-detector --no-extras"
timeout 300 $CMD2 > gpurun_out/${TAG}_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"intersect_wave|interact_wave" -s 24 -c 8 -f -o /tmp/${TAG}_trace $CMD2 > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu trace rc=$?"
python scripts/ncu_summary.py /tmp/${TAG}_trace.ncu-rep gpurun_out/${TAG}_ncu_trace_kernels.txt > /dev/null
timeout 600 ncu --set full --clock-control none -k regex:"pd_field_fast|psf_intensity_kernel|retrace_intersect_wave|gather_segments" -c 10 -f -o /tmp/${TAG}_pd $CMD > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu pd/psf/retrace rc=$?"
python scripts/ncu_summary.py /tmp/${TAG}_pd.ncu-rep gpurun_out/${TAG}_ncu_pd_psf_retrace.txt > /dev/null
timeout 600 ncu --set full --clock-control none -k regex:"compact_fused" -c 8 -f -o /tmp/${TAG}_k3 $CMD2 > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu compaction rc=$?"
python scripts/ncu_summary.py /tmp/${TAG}_k3.ncu-rep gpurun_out/${TAG}_ncu_compaction.txt > /dev/null
ls -la /tmp/${TAG}_*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
