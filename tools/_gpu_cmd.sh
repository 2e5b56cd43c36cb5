set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tests/run_configs.py --config c3 --bytes 1073741824 --check-bytes 268435456 --steps 5 2>&1 | tail -1 | cut -c1-900
PFAC_B200_FILTER=exact timeout 200 python tests/run_configs.py --config c3 --bytes 1073741824 --check-bytes 0 --steps 5 2>&1 | tail -1 | cut -c1-400
timeout 200 python tests/run_configs.py --config c5 --bytes 1073741824 --check-bytes 268435456 --steps 5 2>&1 | tail -1 | cut -c1-900
PFAC_B200_FILTER=exact timeout 200 python tests/run_configs.py --config c5 --bytes 1073741824 --check-bytes 0 --steps 5 2>&1 | tail -1 | cut -c1-400
PFAC_B200_FILTER=hash timeout 200 python tests/run_configs.py --config c2 --check-bytes 268435456 --steps 10 2>&1 | tail -1 | cut -c1-900
timeout 200 python tests/run_configs.py --config c2 --check-bytes 0 --steps 10 2>&1 | tail -1 | cut -c1-400
