#!/bin/sh
# builds the CPU emulation of the kernel core (development check; see hostemu.cpp)
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -O2 -fPIC -shared -std=c++17 -ffp-contract=off -Wall -Wno-unknown-pragmas -Wno-unused-variable -o _build/libhostemu.so hostemu.cpp -lm
