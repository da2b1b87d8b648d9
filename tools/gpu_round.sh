timeout 900 python -m pytest tests -m gpu -x -q -k "compose" 2>&1 | tail -6
python tools/time_compose.py
