#!/bin/bash
# build an A/B variant of libqdx.so with extra nvcc flags:  tools/build_variant.sh <name> <flags...>  -> qdax_b200/libqdx_<name>.so
set -e
name=$1; shift
tmp=$(mktemp -d)
mkdir -p $tmp/qdax_b200 $tmp/include
cp -r qdax_b200/csrc $tmp/qdax_b200/csrc
cp include/qdx.h $tmp/include/
make -C $tmp/qdax_b200/csrc clean >/dev/null
make -C $tmp/qdax_b200/csrc -j4 EXTRA="$*" >/dev/null
cp $tmp/qdax_b200/libqdx.so qdax_b200/libqdx_$name.so
rm -rf $tmp
echo built qdax_b200/libqdx_$name.so
