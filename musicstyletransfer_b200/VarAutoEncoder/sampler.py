"""Inference / style transfer with the reference's interface (VarAutoEncoder/sampler.py): for every target class
the batch's class vector is overwritten, the source is encoded (z = means), and the decoder samples up to 2T
tokens; originals and per-class transfers are written as MIDI files ``out-i.original.mid`` /
``out-i.class-c.mid``.  The autoregressive loop runs on the device (engine.style_transfer)."""
import os
from typing import Optional

import torch

from . import config
from . import model
from . import utils
from .data import Dataset, Loader, MelodyDataset, ToyData
from .utils import to_device_i32
from ..MIDIUtil.Melody import get_melody_from_ids
from ..MIDIUtil.midi_io import MelodyWriter


def load_inference_model(model_folder: str, context, checkpoint: Optional[int], precision="bf16p3f"):
    c = config.Config.load(os.path.join(model_folder, 'config'))
    utils.log_config(c)
    m = model.Model(c, context=context, precision=precision)
    if checkpoint is None:
        return m
    if checkpoint == -1:
        checkpoint = utils.get_latest_checkpoint_index(model_folder)
    utils.load_model_parameters(m, os.path.join(model_folder, 'params.{}'.format(checkpoint)), context)
    return m


def get_sampler(type: str, model_folder: str, context, checkpoint: Optional[int], args):
    if type == 'sampling':
        return Sampling(model_folder, context, checkpoint, verbose=args.verbose,
                        precision=getattr(args, "precision", "tf32"))
    elif type == 'beam-search':
        return BeamSearchSampler(model_folder, context, checkpoint, verbose=args.verbose,
                                 precision=getattr(args, "precision", "tf32"), beam_size=getattr(args, "beam_size", 5))
    raise ValueError("Sampler {} is not implemented".format(type))


class SamplerBase:
    def __init__(self, model_folder: str, context, checkpoint: int, verbose: bool = False, precision="bf16p3f",
                 model_instance=None):
        self.model = model_instance if model_instance is not None else \
            load_inference_model(model_folder, context, checkpoint, precision)
        self.encoder, self.decoder = self.model.encoder, self.model.decoder
        self.model_folder = model_folder
        self.context = context
        self.verbose = verbose
        self.seed = 0

    def reload_checkpoint(self, checkpoint: int):
        self.model = load_inference_model(self.model_folder, self.context, checkpoint)
        self.encoder, self.decoder = self.model.encoder, self.model.decoder

    def update_parameters(self, model):
        self.model = model
        self.encoder, self.decoder = model.encoder, model.decoder

    def _write(self, writer, path, ids):
        writer.write_to_file(path, get_melody_from_ids([int(i) for i in ids]))

    def process_dataset(self, dataset: Dataset, output_suffix: str):
        utils.create_directory_if_not_present(output_suffix)
        print("Starting to decode dataset")
        writer = MelodyWriter()
        current_sample_idx = 0
        for i, batch in enumerate(dataset):
            print("Processing batch {}".format(i))
            self._process(batch, output_suffix, dataset.num_classes(), writer, current_sample_idx)
            current_sample_idx += batch.data[0].shape[0]
        print("Done with dataset decoding")

    def process_batch(self, batch, output_suffix: str, num_classes: int):
        utils.create_directory_if_not_present(output_suffix)
        self._process(batch, output_suffix, num_classes, MelodyWriter(), 0)

    def _process(self, batch, output_suffix, num_classes, writer, base):
        tokens = torch.as_tensor(batch.data[0])
        for i in range(tokens.shape[0]):
            self._write(writer, os.path.join(output_suffix, "out-{}.original.mid".format(base + i)), tokens[i].tolist())
        for class_idx in range(num_classes):
            batch.data[2] = torch.full_like(torch.as_tensor(batch.data[2]), class_idx)     # sampler.py:95 / :126
            sequences = self.sample(batch).cpu()
            for i in range(sequences.shape[0]):
                self._write(writer, os.path.join(output_suffix, "out-{}.class-{}.mid".format(base + i, class_idx)),
                            sequences[i].tolist())

    def read_batch(self, batch):
        dev = self.model.engine.device
        return [to_device_i32(x, dev) for x in batch.data], [to_device_i32(x, dev) for x in batch.label]

    def sample(self, data_batch):
        raise NotImplementedError

    def compute_initial_decoder_state(self, tokens, seq_lens, classes):
        means, _ = self.encoder(tokens, seq_lens, classes)                                   # sampler.py:147-148
        state = self.decoder.get_initial_state(None, classes, means)
        return model.DecoderState(tokens.shape[0], self.model.engine.cfg.dec_layers, state)


class Sampling(SamplerBase):
    def sample(self, data_batch, uniforms=None):
        [tokens, seq_lens, classes], _ = self.read_batch(data_batch)
        if self.verbose:
            print("Inputs to sampling: ")
            print("Tokens: {}, {}".format(tuple(tokens.shape), tokens))
            print("seq_lens: {}, {}".format(tuple(seq_lens.shape), seq_lens))
            print("classes: {}, {}".format(tuple(classes.shape), classes))
        self.seed += 1
        seqs, _ = self.model.engine.style_transfer(tokens, seq_lens, classes, uniforms=uniforms, seed=self.seed)
        return seqs


class BeamSearchSampler(SamplerBase):
    """sampler.py:192-257 on the LSTM decoder (device-side beam search, engine.beam_search).  `sample` returns the best
    hypothesis of every batch row ("take every k-th hypothesis"); `sample_all` returns all beam_size * B of them."""

    def __init__(self, *args, beam_size: int, **kwargs):
        super().__init__(*args, **kwargs)
        self.beam_size = beam_size
        self.max_length_factor = 2.

    def sample_all(self, data_batch):
        [tokens, seq_lens, classes], _ = self.read_batch(data_batch)
        return self.model.engine.beam_search(tokens, seq_lens, classes, self.beam_size)

    def sample(self, data_batch):
        seqs, _ = self.sample_all(data_batch)
        return seqs[::self.beam_size]


def sample_toy(args):
    sampler = get_sampler("sampling", "/tmp/music-style-transfer/toy/model", None, args.checkpoint, args)
    sampler.process_dataset(ToyData(), args.out_samples)


def main(argv=None):
    args = config.get_config(argv)
    if args.toy:
        sample_toy(args)
        return
    loader = Loader(path=args.data, max_sequence_length=args.max_seq_len,
                    slices_per_quarter_note=args.slices_per_quarter_note)
    dataset = MelodyDataset(args.batch_size, loader.max_sequence_length, loader.melodies)
    sampler = get_sampler(args.sampling_type, args.model_output, None, args.checkpoint, args)
    sampler.process_dataset(dataset, args.out_samples)


if __name__ == '__main__':
    main()
