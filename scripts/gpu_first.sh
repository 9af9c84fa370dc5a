set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as ge; ge.smoke()" 2>&1 | tail -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
