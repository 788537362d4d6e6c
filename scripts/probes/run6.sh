set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "strips or multigpu or c4 or mirror or fuzz" 2>&1 | tail -5
python scripts/strips_model.py 2>&1 | tail -10
