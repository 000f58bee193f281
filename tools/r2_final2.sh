#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q -k "awkward or options or lattice_through or tiny or incremental" 2>&1 | tail -3
