#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gridbench.py
for n in "$@"; do XNART_LIB=$PWD/build/variants/libxnart_$n.so python tools/gridbench.py; done
