# launch list + ncu --set full of selected kernels of the bench.  Usage: gpurun -- bash scripts/gpu_prof.sh TAG "regex" [skip] [count]
TAG=${1:-prof}; KRE=${2:-intersect_wave}; SKIP=${3:-4}; CNT=${4:-4}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/${TAG}*
