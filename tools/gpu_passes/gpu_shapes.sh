#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_launch_shapes.py -q -m gpu 2>&1 | tail -6
