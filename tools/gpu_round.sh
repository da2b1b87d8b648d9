mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "skybox_synth" 2>&1 | tail -25
