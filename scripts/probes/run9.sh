mkdir -p gpurun_out
for f in 1 4 8; do
echo "== default grid fpc=$f"; python scripts/strips_model.py $f 2>&1 | grep -E "^8 |^1 busy|^4 busy"
echo "== GRB_LIST_GRID=1480 fpc=$f"; GRB_LIST_GRID=1480 python scripts/strips_model.py $f 2>&1 | grep -E "^8 |^4 busy"
done
