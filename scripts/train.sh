#!/bin/bash
# The reference's scripts/train.sh:5-23 launches `music_style_transfer.GAN.main`, a package that is not part of
# the reference tree.  Its flag set is kept (unknown --g-*/--d-*/--noise-dim/--discriminator-update-steps flags
# are ignored by parse_known_args, config.py:74) and mapped onto the VarAutoEncoder style-transfer path
# (SURVEY.md §2 row 21): encode the source, swap the class, decode (sampler.py:77-135).
cd "$(dirname "$0")/.." || exit 1
python -m music_style_transfer.VarAutoEncoder.sampler \
--batch-size 32 \
--out-samples /tmp/out \
--max-seq-len 64 \
--slices-per-quarter-note 4 \
--data ${MSX_DATA:-./work/data/guitar_bass} \
--sampling-frequency 50 \
--epochs 10000 \
--discriminator-update-steps 5 \
--model-output ${MSX_MODEL:-models/guitar_bass} \
--g-learning-rate 0.00005 \
--g-n-layers 1 \
--g-rnn-hidden-dim 256 \
--g-emb-hidden-dim 256 \
--noise-dim 64 \
--d-learning-rate 0.00005 \
--d-n-layers 1 \
--d-rnn-hidden-dim 256 \
--d-emb-hidden-dim 256 --gpu "$@"
