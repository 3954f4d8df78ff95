#!/bin/bash
# Real-jar validation kit, part 2 (needs a JDK >= 13; this repository's build image has none):
#   run the UNMODIFIED reference on the kit's FASTQ, time it, then diff its read names against the GPU path (or the CPU oracle).
#
#   baseline/run_java_baseline.sh <kit_dir> <path to /root/reference/Jar> [bcEditDistance] [threads]
#
# Uses `scanfastq -g <list>` so that the second pass runs against exactly the kit's barcode list (README.md:214-226: "the first pass will be
# skipped and the supplied barcode list will be used").  The jar looks for config.xml in the working directory, then next to itself.
set -euo pipefail
kit=${1:?kit dir (baseline/make_fastq_kit.py)}; jar=${2:?directory of NanoporeBC_UMI_finder-2.1.jar}; ed=${3:-2}; thr=${4:-$(nproc)}
command -v java >/dev/null || { echo "no java on PATH: run this on a box with a JDK >= 13"; exit 2; }
out=$kit/scan_ed$ed; rm -rf "$out"; mkdir -p "$out"
start=$(date +%s.%N)
( cd "$jar" && java -Xmx16g -jar NanoporeBC_UMI_finder-2.1.jar scanfastq -d "$kit/fastq_pass" -o "$out" --bcEditDistance "$ed" -g "$kit/barcodes.tsv" -t "$thr" ) | tee "$out/scanfastq.log"
end=$(date +%s.%N)
n=$(( $(wc -l < "$kit/fastq_pass/reads.fastq") / 4 ))
echo "JVM scanfastq: $n reads in $(echo "$end - $start" | bc) s on $thr threads (whole step: chimera split + polyA/adapter scan + barcode assignment + FASTQ I/O;"
echo "the 'Barcode Search' line of the CpuTimeStats block in $out/scanfastq.log is the barcode-assignment share)"
here=$(cd "$(dirname "$0")/.." && pwd)
python "$here/baseline/compare_with_jar.py" --scan-dir "$out" --list "$kit/barcodes.tsv" --ed "$ed" "${@:5}"
