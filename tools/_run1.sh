timeout 600 python -m pytest tests -m gpu -x -q -k "polar_lean" 2>&1 | tail -5
