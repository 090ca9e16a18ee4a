#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_power_of_two.py -q -m gpu 2>&1 | tail -15
