bash tools/gpu_round.sh u20
bash tools/gpu_profile.sh r1 55
