timeout 900 python -m pytest tests -m gpu -x -q -k "skybox" 2>&1 | tail -3
python tools/time_sky.py 2048 2048
