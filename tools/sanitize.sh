#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the engine's kernel bodies and host code: the emulated build (csrc compiled
# for the host with -DABM_EMU, tests/emu) is rebuilt with -fsanitize=address,undefined, put in the emulated library's place, and
# the CPU parity / strips / host-API tests and the fuzzers run on it (SURVEY.md §5 "race detection / sanitizers").  The plain
# emulated library is restored afterwards.  Any report aborts the run (-fno-sanitize-recover).
#   bash tools/sanitize.sh [fuzz seeds, default 300]
set -e
cd "$(dirname "$0")/.."
SEEDS=${1:-300}
EMU=tests/emu/libabm_emu.so
python -c "import abm_b200; from abm_b200 import build; build.build_emu(); build.build_oracle()"
cp $EMU /tmp/libabm_emu_plain.so
trap 'cp /tmp/libabm_emu_plain.so '$EMU'; touch '$EMU EXIT
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=undefined \
    -Wno-unknown-pragmas -ffp-contract=off -DABM_EMU -x c++ -shared -o $EMU animal-movement-abm_b200/csrc/abm_engine.cu 2>/dev/null
touch $EMU
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1
python -m pytest tests/test_parity.py tests/test_strips.py tests/test_host_api.py tests/test_abi.py -x -q -m "not gpu" -p no:cacheprovider
python tools/fuzz.py engine 930000 $SEEDS
python tools/fuzz.py astar 940000 $SEEDS
FUZZ_SCHED=1 ABM_TILE_ROUNDS=1 FORCE_GENERAL=1 python tools/fuzz.py engine 950000 $SEEDS
# the oracle itself under the sanitizers (its own tests and the example scripts)
make -C oracle asan >/dev/null
cp oracle/liboracle.so /tmp/liboracle_plain.so
cp oracle/liboracle_asan.so oracle/liboracle.so; touch oracle/liboracle.so
python -m pytest tests/test_oracle.py tests/test_examples.py -x -q -p no:cacheprovider || { cp /tmp/liboracle_plain.so oracle/liboracle.so; exit 1; }
cp /tmp/liboracle_plain.so oracle/liboracle.so; touch oracle/liboracle.so; rm -f oracle/liboracle_asan.so
# second pass: the threads of every CTA phase visited backwards (-DABM_EMU_REVERSE): a result that depends on the order in which the
# threads of a phase run would be a race between barriers on the device
unset LD_PRELOAD
g++ -O2 -g -std=c++17 -fPIC -Wno-unknown-pragmas -ffp-contract=off -DABM_EMU -DABM_EMU_REVERSE -x c++ -shared -o $EMU animal-movement-abm_b200/csrc/abm_engine.cu 2>/dev/null
touch $EMU
python -m pytest tests/test_parity.py tests/test_strips.py -x -q -m "not gpu" -p no:cacheprovider
python tools/fuzz.py engine 960000 $SEEDS
python tools/fuzz.py astar 970000 $SEEDS
FUZZ_SCHED=1 ABM_TILE_ROUNDS=1 python tools/fuzz.py engine 980000 $SEEDS
