#!/usr/bin/env bash
# Builds the drop-in `od-msspe` executable (C++ host over the C ABI) into open-msspe-design_b200/bin/.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
mkdir -p "$PKG/bin"
g++ -O2 -std=c++17 -Wall -ffp-contract=off -o "$PKG/bin/od-msspe" "$HERE/od_msspe_main.cpp" \
  -L"$PKG" -lodmsspe_b200 -Wl,-rpath,'$ORIGIN/..' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $PKG/bin/od-msspe"
