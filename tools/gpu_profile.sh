#!/bin/bash
# One GPU-box visit that collects the ncu evidence for profiles/: launch lists of one generator step and of the encoder path, and
# --set full captures (raw + source pages exported as CSV on the box) of the dominant kernels.  Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
O=gpurun_out/$TAG
mkdir -p $O
# launch list of one batch-8 generator step (tools/prof_one.py: 1 warm-up step + 1 measured step; the second half of the list)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c2.csv python tools/prof_one.py 2 > $O/launches_c2.log 2>&1
# launch list of the encoder path (encode + 2 x AR_eval_forward)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_enc.csv python tools/prof_encoder.py > $O/launches_enc.log 2>&1
# full captures: the convolution launches of the second step (skip the first step's), the renderer, the FIR epilogue
NCONV=$(grep -c "conv_tc" $O/launches_c2.csv)
ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s $((NCONV / 2)) -c 60 -o $O/conv_full python tools/prof_one.py 2 > $O/conv_full.log 2>&1
ncu -i $O/conv_full.ncu-rep --page raw --csv > $O/conv_raw.csv 2>/dev/null
rm -f $O/conv_full.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o $O/render_full python tools/prof_one.py 2 > $O/render_full.log 2>&1
ncu -i $O/render_full.ncu-rep --page raw --csv > $O/render_raw.csv 2>/dev/null
ncu -i $O/render_full.ncu-rep --page source --csv | gzip > $O/render_source.csv.gz
# the reports themselves are too large for the 64 MiB return channel: the CSV exports above are what travels
rm -f $O/*.ncu-rep $O/*.ncu-rep.tmp
du -sh $O
ls -la $O
