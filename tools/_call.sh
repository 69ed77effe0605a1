timeout 600 python -m pytest tests/test_gpu_wire_formats.py -x -q 2>&1 | tail -8
