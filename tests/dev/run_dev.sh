python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python tests/dev/dev_tile.py c5 65536 100 2>&1 | tail -3
