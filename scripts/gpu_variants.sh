for it in 1 2; do
echo inline $it; python scripts/vertical_timeline.py --cfg unsat_inline_iters=$it 2>&1 | tail -3
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
