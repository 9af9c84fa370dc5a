TAG=${1:-pd}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
python scripts/pd_speed.py 32 1024 2>&1 | tail -3
python scripts/pd_speed.py 64 2048 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(json.dumps(d['detector'], indent=1)); print('C2 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']))"
