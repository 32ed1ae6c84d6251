#!/bin/bash
# run tools/pathbench.py quick for the default library and every build/variants/libxnart_*.so named on the command line
cd "$(dirname "$0")/.."
python tools/pathbench.py quick
for n in "$@"; do XNART_LIB=$PWD/build/variants/libxnart_$n.so python tools/pathbench.py quick; done
