timeout 900 python -m pytest tests -m gpu -x -q -k "stream_ordering" 2>&1 | tail -8
