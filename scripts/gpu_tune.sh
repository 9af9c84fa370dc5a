# parity + quick bench for each value of a tuning env var.  Usage: gpurun -- bash scripts/gpu_tune.sh VAR "v1 v2 ..." TAG
VAR=$1; VALS=$2; TAG=${3:-tune}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_$v.err | tee gpurun_out/${TAG}_$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g e2e %.4g ms/step %.3f  k1 ms/launch %.4f frac %.4f sdf/step %.4g launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['counted']['sdf_evals_per_step'], d['gpu_launches']))"
done
