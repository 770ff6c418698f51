#!/bin/bash
# the host-side readers and the packer (dmx_bam.cpp, dmx_vcf.cpp, dmx_bgzf.h, dmx_pack.cpp) under AddressSanitizer + UBSan and under ThreadSanitizer:
# every input form, whole and damaged, default windows and windows of a few bytes, several threads.  No GPU needed.
# usage: bash tools/san_readers.sh [outfile]
set -e
HERE=$(cd $(dirname $0)/.. && pwd)
C=$HERE/demuxlet_b200/csrc
D=$(mktemp -d)
python $HERE/tools/san_inputs.py $D
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-omit-frame-pointer -o $D/drv_asan $HERE/tools/san_driver.cpp $C/dmx_bam.cpp $C/dmx_vcf.cpp $C/dmx_pack.cpp -lz -lpthread
g++ -O1 -g -std=c++17 -fsanitize=thread -o $D/drv_tsan $HERE/tools/san_driver.cpp $C/dmx_bam.cpp $C/dmx_vcf.cpp $C/dmx_pack.cpp -lz -lpthread
OUT=$(realpath -m ${1:-$D/report.txt})
cd $D
export DMX_BAM_MIN_CHUNK=64
{
for drv in drv_asan drv_tsan; do
  for cfg in "X=1" "DMX_IO_BATCH_BLOCKS=2 DMX_IO_CHUNK_BYTES=999" "DMX_IO_BATCH_BLOCKS=1 DMX_IO_CHUNK_BYTES=13"; do
    echo "=== $drv $cfg"
    for f in a.bam b.bam c.sam bad.bam cut.bam garbage.bam; do
      echo "-- $f"; env $cfg DMX_BAM_THREADS=4 ./$drv bam $f snps.bin 2>&1 | grep -E "ERROR|WARNING|runtime error|SUMMARY|^bam|^pack" | head -6
    done
    for f in GP.vcf GP.vcf.gz GP.bcf PL.vcf.gz PL.bcf GP_cut.bcf garbage.vcf; do
      echo "-- $f"; env $cfg DMX_VCF_THREADS=4 ./$drv vcf $f 2>&1 | grep -E "ERROR|WARNING|runtime error|SUMMARY|^vcf" | head -6
    done
  done
done
} > $OUT.tmp 2>&1 || true
N=$(grep -c -E "ERROR|WARNING: ThreadSanitizer|runtime error" $OUT.tmp || true)
{ echo "sanitizer reports: $N"; cat $OUT.tmp; } > $OUT
rm -f $OUT.tmp
[ -n "$1" ] || cat $OUT
rm -rf $D
