"""Training loop (mirrors helen/modules/python/models/train.py:19-259): epochs over the training images, one optimisation
step per chunk, evaluation and a checkpoint after every epoch.  The per-chunk body (:189-201: forward, the two
criteria, backward) is `ChunkTrainer.step` -> hb_train_step_chunk; Adam and ReduceLROnPlateau are torch's, untouched.
GPU only.  One process per GPU: with a `dist_ctx` (models/train_distributed.py) each rank trains on its share of the
images and the ranks' gradients are averaged after every step (models/grad_sync.py), which is what the reference's
DistributedDataParallel wrapper does (train_distributed.py:128-131); without one this is the single-GPU loop."""
import os
import sys

import torch
from torch.utils.data import DataLoader
from torch.utils.data.distributed import DistributedSampler

from ..options import ImageSizeOptions, TrainOptions
from ..TextColor import TextColor
from .dataloader import SequenceDataset
from .grad_sync import FlatGradients
from .ModelHander import ModelHandler
from .test import test
from .train_step import ChunkTrainer


def train(train_file, test_file, batch_size, epoch_limit, gpu_mode, num_workers, retrain_model, retrain_model_path,
          gru_layers, hidden_size, lr, decay, model_dir, stats_dir, not_hyperband, dist_ctx=None):
    is_main = dist_ctx is None or dist_ctx.is_main          # rank 0 logs, evaluates and saves (train_distributed.py:121-124)
    if not gpu_mode:
        sys.stderr.write(TextColor.RED + "ERROR: helen_b200 HAS NO CPU PATH, USE gpu_mode.\n" + TextColor.END)
        exit(1)
    write_logs = (not_hyperband is True) and is_main
    train_loss_logger = open(stats_dir + "train_loss.csv", 'w') if write_logs else None
    test_loss_logger = open(stats_dir + "test_loss.csv", 'w') if write_logs else None
    confusion_matrix_logger = open(stats_dir + "confusion_matrix.txt", 'w') if write_logs else None

    sys.stderr.write(TextColor.PURPLE + 'Loading data\n' + TextColor.END)
    train_data_set = SequenceDataset(train_file)
    sampler = None if dist_ctx is None else DistributedSampler(train_data_set, num_replicas=dist_ctx.world_size, rank=dist_ctx.rank)
    train_loader = DataLoader(train_data_set, batch_size=batch_size, shuffle=sampler is None, sampler=sampler,
                              num_workers=num_workers, pin_memory=True)
    image_features = int(train_data_set[0][0].shape[1]) if len(train_data_set) else ImageSizeOptions.IMAGE_HEIGHT
    if retrain_model is True:                                                # train.py:78-95
        if os.path.isfile(retrain_model_path) is False:
            sys.stderr.write(TextColor.RED + "ERROR: INVALID PATH TO RETRAIN PATH MODEL --retrain_model_path\n")
            exit(1)
        transducer_model, hidden_size, gru_layers, prev_ite = ModelHandler.load_simple_model(
            retrain_model_path, input_channels=ImageSizeOptions.IMAGE_CHANNELS, image_features=image_features,
            seq_len=ImageSizeOptions.SEQ_LENGTH, num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS,
            num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS)
        if not_hyperband is True:
            epoch_limit = prev_ite + epoch_limit
    else:
        transducer_model = ModelHandler.get_new_gru_model(
            input_channels=ImageSizeOptions.IMAGE_CHANNELS, image_features=image_features, gru_layers=gru_layers,
            hidden_size=hidden_size, num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS,
            num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS)
        prev_ite = 0
    transducer_model = transducer_model.cuda()
    if dist_ctx is not None:
        dist_ctx.broadcast_parameters(transducer_model)
    trainer = ChunkTrainer(transducer_model, class_weights=TrainOptions.CLASS_WEIGHTS,      # train.py:121-126
                           device=torch.cuda.current_device())
    shared_grads = FlatGradients(transducer_model.parameters()) if dist_ctx is not None else None
    param_count = sum(p.numel() for p in transducer_model.parameters())
    if is_main:
        sys.stderr.write(TextColor.RED + "INFO: TOTAL TRAINABLE PARAMETERS:\t" + str(param_count) + "\n" + TextColor.END)
    model_optimizer = torch.optim.Adam(transducer_model.parameters(), lr=lr, weight_decay=decay)          # :110
    lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(model_optimizer, 'min')                      # :112
    if retrain_model is True:
        checkpoint = ModelHandler.load_checkpoint(retrain_model_path)
        if checkpoint.get('model_optimizer'):
            model_optimizer.load_state_dict(checkpoint['model_optimizer'])

    stats = {'loss_epoch': [], 'accuracy_epoch': []}
    sys.stderr.write(TextColor.PURPLE + 'Training starting\n' + TextColor.END)
    for epoch in range(prev_ite, epoch_limit, 1):
        total_loss_base = total_loss_rle = total_loss = 0.0
        total_images, batch_no = 0, 1
        if is_main:
            sys.stderr.write(TextColor.BLUE + 'Train epoch: ' + str(epoch + 1) + "\n")
        if sampler is not None:
            sampler.set_epoch(epoch)
        for images, label_base, label_rle in train_loader:
            images = images.cuda(non_blocking=True).float()
            label_base, label_rle = label_base.cuda(non_blocking=True).long(), label_rle.cuda(non_blocking=True).long()
            hidden = torch.zeros(images.size(0), 2 * TrainOptions.GRU_LAYERS, TrainOptions.HIDDEN_SIZE, device="cuda")
            for i in range(0, images.size(1), TrainOptions.WINDOW_JUMP):                      # :174
                if shared_grads is None:
                    model_optimizer.zero_grad()
                if i + TrainOptions.TRAIN_WINDOW > images.size(1):
                    break
                loss, loss_base, loss_rle, hidden = trainer.step(                                  # :189-201, :206
                    images[:, i:i + TrainOptions.TRAIN_WINDOW], hidden,
                    label_base[:, i:i + TrainOptions.TRAIN_WINDOW], label_rle[:, i:i + TrainOptions.TRAIN_WINDOW])
                if shared_grads is not None:
                    shared_grads.all_reduce_mean(dist_ctx.group)      # the step overwrote every p.grad (views of one buffer)
                model_optimizer.step()                                                             # :202
                total_loss += loss
                total_loss_base += loss_base
                total_loss_rle += loss_rle
                total_images += images.size(0)
            avg_loss = (total_loss / total_images) if total_images else 0
            if write_logs:
                train_loss_logger.write(str(epoch + 1) + "," + str(batch_no) + "," + str(avg_loss) + "\n")
            batch_no += 1
        if is_main:
            sys.stderr.write("Base: " + str(round(total_loss_base, 4)) + ", RLE: " + str(round(total_loss_rle, 4))
                             + ", TOTAL: " + str(round(total_loss, 4)) + "\n")

        if dist_ctx is not None:
            dist_ctx.assert_replicas_identical(transducer_model)
            dist_ctx.barrier()                                # train_distributed.py:243
        if is_main:
            stats_dictionary = test(test_file, batch_size, gpu_mode, transducer_model, num_workers, gru_layers, hidden_size,
                                    num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS, num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS)
        else:
            stats_dictionary = {'loss': 0.0, 'accuracy': 0}
        if dist_ctx is not None:                              # every rank's scheduler sees rank 0's evaluation loss
            stats_dictionary['loss'] = dist_ctx.broadcast_value(stats_dictionary['loss'], torch.device("cuda", torch.cuda.current_device()))
        stats['loss'] = stats_dictionary['loss']
        stats['accuracy'] = stats_dictionary['accuracy']
        stats['train_loss'] = total_loss
        stats['loss_epoch'].append((epoch, stats_dictionary['loss']))
        stats['accuracy_epoch'].append((epoch, stats_dictionary['accuracy']))
        lr_scheduler.step(stats['loss'])
        if write_logs:
            ModelHandler.save_model(transducer_model, model_optimizer, hidden_size, gru_layers, epoch,
                                    model_dir + "HELEN_epoch_" + str(epoch + 1) + '_checkpoint.pkl')
            sys.stderr.write(TextColor.RED + "\nMODEL SAVED SUCCESSFULLY.\n" + TextColor.END)
            test_loss_logger.write(str(epoch + 1) + "," + str(stats['loss']) + "," + str(stats['accuracy']) + "\n")
            confusion_matrix_logger.write(str(epoch + 1) + "\n" + str(stats_dictionary['base_confusion_matrix']) + "\n")
            train_loss_logger.flush()
            test_loss_logger.flush()
            confusion_matrix_logger.flush()
    trainer.close()
    if is_main:
        sys.stderr.write(TextColor.PURPLE + 'Finished training\n' + TextColor.END)
    return transducer_model, model_optimizer, stats
