#!/bin/bash
# Builds kernel variants into variants/<name>.so (experiments; see csrc/lovasz.cu RTDA_* macros).
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() { # name flags...
  name=$1; shift
  make -s -C rtda_semanticsegmentation_b200/csrc OUT=../../variants/$name.so EXTRA="$*" -B
  echo built variants/$name.so
}
"$@"
