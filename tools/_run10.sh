timeout 100 python tools/march_check.py check quick 2>&1 | tail -1
for lag in 0 12 24 32; do echo "== LAG_U=$lag"; TTCR_B200_LAG_U=$lag timeout 100 python tools/march_check.py time 512 2>&1 | grep '"kernel": 7' | grep -v depth; done
