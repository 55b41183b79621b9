#!/bin/bash
# The device headers (fe / sc / sha512 / ge / verify_core, host-compiled by tests/hostemu) under UndefinedBehaviorSanitizer: every
# hostemu test must pass with -fno-sanitize-recover (shifts, signed overflow, misaligned or out-of-range accesses abort).  ~3.5 min.
set -e
cd "$(dirname "$0")/.."
python -c "from hotstuff_b200 import build; build.build_hostemu()"
cp tests/hostemu/libhs_hostemu.so /tmp/libhs_hostemu_plain.so
g++ -O1 -g -std=c++17 -fPIC -shared -DHS_HOST_EMU -Wno-unknown-pragmas -fsanitize=undefined -fno-sanitize-recover=undefined \
    -o tests/hostemu/libhs_hostemu.so tests/hostemu/hostemu.cpp
trap 'cp /tmp/libhs_hostemu_plain.so tests/hostemu/libhs_hostemu.so' EXIT
python -m pytest tests/test_hostemu.py tests/test_hostemu_properties.py -x -q
