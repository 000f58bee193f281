#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
