# The CPU tests (oracle KATs, golden fixtures, host parity) against the oracle built with UndefinedBehaviorSanitizer.
# Usage: bash scripts/oracle_ubsan.sh      (no GPU needed; any undefined behaviour aborts the test process)
set -e
cd "$(dirname "$0")/.."
make -C oracle ubsan
ORACLE_LIB=$PWD/oracle/liboracle_ubsan.so UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 python -m pytest tests -q -m "not gpu" -k "oracle or kat or golden or host_parity or psf or polfilter or asphere or cylindrical or retrace"
