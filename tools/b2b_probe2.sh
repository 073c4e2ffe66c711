for envs in "" "INFUR_B200_NO_B2B=1" "INFUR_B200_NO_GRAPH=1" "INFUR_B200_NO_GRAPH=1 INFUR_B200_NO_B2B=1" "INFUR_B200_NO_PDL=1" "INFUR_B200_NO_PDL=1 INFUR_B200_NO_B2B=1"; do
  echo "== $envs"; env $envs python tools/profile_step.py --iters 0 --steps 30 2>&1 | tail -1
done
