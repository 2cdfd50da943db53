#!/bin/bash
python -m pytest tests/test_gpu_label.py -x -q 2>&1 | tail -12
