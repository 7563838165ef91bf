"""Drop-in package name of the reference (slyforce/MusicStyleTransfer): every module re-exports the
B200-native implementation in ``musicstyletransfer_b200`` under the reference's module path, so
``python -m music_style_transfer.VarAutoEncoder.main <flags>`` (scripts/train-vae.sh) runs unchanged."""
