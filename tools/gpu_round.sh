timeout 900 python -m pytest tests -m gpu -x -q -k "shard_invariant" 2>&1 | tail -6
