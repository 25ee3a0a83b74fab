#!/bin/bash
# builds a kernel-experiment variant of the library next to the product one: tools/build_variant.sh NAME "-DFLAG=..."
# -> skirt_b200/variants/libskirtgpu_NAME.so (selected at run time with SKG_LIBRARY, see skirt_b200/binding.py)
set -e
name=$1; shift
mkdir -p skirt_b200/variants
make -j8 EXTRA="$*" BUILD=/tmp/skg_variant_$name LIB=skirt_b200/variants/libskirtgpu_$name.so skirt_b200/variants/libskirtgpu_$name.so > /tmp/skg_variant_$name.log 2>&1 || { tail -20 /tmp/skg_variant_$name.log; exit 1; }
grep -A2 "absorbStageILi0ELb1ELb1\|peelStageILi0ELb1\|propagateStageILi0ELb1" /tmp/skg_variant_$name/mc_kernels.ptxas.log | grep "Used" | tr '\n' ' '; echo " <- $name"
