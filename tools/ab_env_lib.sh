#!/bin/bash
# tools/ab_env_lib.sh <variant> "NAME=VAL ..." ... : environment A/B on tools/libpagk_<variant>.so
cd "$(dirname "$0")/.."
export PAGK_LIB=$PWD/tools/libpagk_$1.so; shift
exec tools/ab_env.sh "$@"
