#!/bin/bash
# Profiling build of the BiLSTM kernels (clock64 phase counters printed by the kernels), then the normal build again.
set -e
cd "$(dirname "$0")/.."
VOCR_NVCC_FLAGS=-DVOCR_LSTM_PROF python -c "from vistaocr_b200.build import build; build(force=True)" >/dev/null
ITERS=1 timeout 200 python tools/lstm_bench.py
python -c "from vistaocr_b200.build import build; build(force=True)" >/dev/null
