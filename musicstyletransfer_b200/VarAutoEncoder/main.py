"""Training entry point with the reference's command line (VarAutoEncoder/main.py, scripts/train-vae.sh).

    python -m music_style_transfer.VarAutoEncoder.main <flags of config.py>

Under ``torchrun`` (one process per GPU) the same command trains data-parallel: every rank builds the same
dataset order, takes rows rank::world of each batch and the gradient arena is all-reduced over NCCL."""
import os

import torch

from . import model, trainer
from .config import get_config
from .data import Loader, RollDataset, ToyData, load_dataset
from .sampler import Sampling
from .transformer import TransformerConfig
from .utils import create_directory_if_not_present, log_config, log_model_variables


def create_toy_model_config(data):
    """main.py:14-38."""
    tc = lambda: TransformerConfig(model_size=32, dropout=0.0, num_layers=1, vocab_size=data.num_tokens(), num_heads=2)
    return model.ModelConfig(
        encoder_config=model.EncoderConfig(transformer_config=tc(), latent_dim=16, num_classes=data.num_classes(),
                                           input_dim=data.num_tokens()),
        decoder_config=model.DecoderConfig(transformer_config=tc(), latent_dim=16, num_classes=data.num_classes(),
                                           output_dim=data.num_tokens()))


def create_toy_train_config():
    """main.py:41-55."""
    return trainer.TrainConfig(batch_size=1, sampling_frequency=500, checkpoint_frequency=1000,
                               num_checkpoints_not_improved=-1, kl_loss=1.0,
                               optimizer=trainer.OptimizerConfig(learning_rate=1e-3, optimizer='adam',
                                                                 optimizer_params='clip_gradient:1.0'),
                               label_smoothing=0.0, negative_label_downscaling=True, verbose=False)


def main_toy(epochs=20000, model_folder="/tmp/music-style-transfer/toy/model", precision="fp32", max_steps=-1):
    dataset = ToyData()
    cfg = create_toy_model_config(dataset)
    m = model.Model(config=cfg, precision=precision)
    create_directory_if_not_present(model_folder)
    cfg.save(os.path.join(model_folder, 'config'))
    t = trainer.Trainer(config=create_toy_train_config(), context=None, model=m, sampler=None, max_steps=max_steps)
    t.fit(dataset=dataset, validation_dataset=dataset, model_folder=model_folder, epochs=epochs)
    return t


def create_train_config(args):
    return trainer.TrainConfig(batch_size=args.batch_size, sampling_frequency=args.sampling_frequency,
                               checkpoint_frequency=args.checkpoint_frequency,
                               num_checkpoints_not_improved=args.num_checkpoints_not_improved, kl_loss=args.kl_loss,
                               optimizer=trainer.OptimizerConfig(learning_rate=args.learning_rate,
                                                                 optimizer=args.optimizer,
                                                                 optimizer_params=args.optimizer_params),
                               label_smoothing=args.label_smoothing,
                               negative_label_downscaling=args.negative_label_downscaling, verbose=args.verbose)


def create_model_config(args, dataset):
    """main.py:96-118.  --decoder-type lstm builds the LSTM decoder the --d-* flags describe; transformer builds
    the HEAD ``Decoder`` with model_size = --d-rnn-hidden-dim, layers = --d-n-layers, heads = --e-num-heads."""
    enc = model.EncoderConfig(
        transformer_config=TransformerConfig(model_size=args.e_rnn_hidden_dim, dropout=args.e_dropout,
                                             num_layers=args.e_n_layers, vocab_size=dataset.num_tokens(),
                                             num_heads=args.e_num_heads),
        latent_dim=args.latent_dim, num_classes=dataset.num_classes(), input_dim=dataset.num_tokens())
    if args.decoder_type == "lstm":
        dec = model.DecoderConfig(lstm_config=model.LSTMConfig(n_layers=args.d_n_layers, hidden_dim=args.d_rnn_hidden_dim,
                                                               dropout=args.d_dropout),
                                  latent_dim=args.latent_dim, num_classes=dataset.num_classes(),
                                  output_dim=dataset.num_tokens())
    else:
        dec = model.DecoderConfig(transformer_config=TransformerConfig(model_size=args.d_rnn_hidden_dim,
                                                                       dropout=args.d_dropout,
                                                                       num_layers=args.d_n_layers,
                                                                       vocab_size=dataset.num_tokens(),
                                                                       num_heads=args.e_num_heads),
                                  latent_dim=args.latent_dim, num_classes=dataset.num_classes(),
                                  output_dim=dataset.num_tokens())
    return model.ModelConfig(encoder_config=enc, decoder_config=dec)


def _init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world


def main(argv=None):
    args = get_config(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("this implementation has no CPU context: a CUDA device (B200, sm_100a) is required")
    _init_distributed()
    if args.toy:
        main_toy(precision=args.precision, max_steps=args.max_steps)
        return
    loader = Loader(path=args.data, max_sequence_length=args.max_seq_len,
                    slices_per_quarter_note=args.slices_per_quarter_note)
    val_loader = None
    if args.validation_data is not None:
        val_loader = Loader(path=args.validation_data, max_sequence_length=args.max_seq_len,
                            slices_per_quarter_note=args.slices_per_quarter_note)
    roll_mode = getattr(args, "featurisation", "events") == "roll"
    if roll_mode:
        assert args.decoder_type == "lstm", "--featurisation roll uses the LSTM decoder"
        train_dataset = RollDataset(args.batch_size, args.max_seq_len, loader.melodies, args.slices_per_quarter_note)
        valid_dataset = (RollDataset(args.batch_size, args.max_seq_len, val_loader.melodies, args.slices_per_quarter_note)
                         if val_loader is not None else None)
    else:
        train_dataset, valid_dataset = load_dataset(loader, args.batch_size, args.validation_split, val_loader,
                                                    device_rows=getattr(args, "device_dataset", True))
    create_directory_if_not_present(args.model_output)
    if args.out_samples:
        create_directory_if_not_present(args.out_samples)
    cfg = create_model_config(args, train_dataset)
    cfg.save(args.model_output + '/config')
    log_config(cfg)
    m = model.Model(config=cfg, precision=args.precision, seed=args.seed, featurisation="roll" if roll_mode else "events")
    log_model_variables(m)
    # sampling decodes token events; the piano-roll path trains / validates only
    sampler = None if roll_mode else Sampling(args.model_output, None, None, verbose=args.verbose, model_instance=m)
    t = trainer.Trainer(config=create_train_config(args), context=None, model=m, sampler=sampler, log_dir=args.log_dir,
                        max_steps=args.max_steps, cuda_graph=getattr(args, "cuda_graph", True))
    t.fit(dataset=train_dataset, validation_dataset=valid_dataset, model_folder=args.model_output, epochs=args.epochs)
    print("Training finished.")


if __name__ == '__main__':
    main()
