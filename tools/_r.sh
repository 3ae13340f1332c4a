bash tools/gpu_round_check.sh v17
