"""Evaluation pass (mirrors helen/modules/python/models/test.py:14-167): sliding-window forward over every test image,
the two cross-entropy losses and the base / run-length confusion matrices.  The forward runs through the CUDA library
(`TransducerGRU.forward` -> hb_forward_chunk); the confusion matrices are counted with torch.bincount (the reference uses
torchnet's ConfusionMeter, same [target, prediction] layout).  GPU only."""
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ..options import ImageSizeOptions, TrainOptions
from ..TextColor import TextColor
from .dataloader import SequenceDataset


def _confusion(logits, labels, classes):
    pred = logits.reshape(-1, classes).argmax(1)
    return torch.bincount(labels.reshape(-1) * classes + pred, minlength=classes * classes).reshape(classes, classes)


def test(data_file, batch_size, gpu_mode, transducer_model, num_workers, gru_layers, hidden_size,
         num_base_classes=ImageSizeOptions.TOTAL_BASE_LABELS, num_rle_classes=ImageSizeOptions.TOTAL_RLE_LABELS):
    if not gpu_mode:
        sys.stderr.write(TextColor.RED + "ERROR: helen_b200 HAS NO CPU PATH, USE gpu_mode.\n" + TextColor.END)
        exit(1)
    test_loader = DataLoader(SequenceDataset(data_file), batch_size=batch_size, shuffle=False, num_workers=num_workers, pin_memory=True)
    class_weights = torch.Tensor(TrainOptions.CLASS_WEIGHTS).cuda()
    criterion_base = nn.CrossEntropyLoss()                                   # test.py:52-54
    criterion_rle = nn.CrossEntropyLoss(weight=class_weights)
    sys.stderr.write(TextColor.PURPLE + 'Test starting\n' + TextColor.END)
    base_cm = torch.zeros(num_base_classes, num_base_classes, dtype=torch.int64, device="cuda")
    rle_cm = torch.zeros(num_rle_classes, num_rle_classes, dtype=torch.int64, device="cuda")
    total_loss, total_loss_rle, total_images = 0.0, 0.0, 0
    with torch.no_grad():
        for images, label_base, label_rle in test_loader:
            images = images.cuda().float()
            label_base, label_rle = label_base.cuda().long(), label_rle.cuda().long()
            hidden = torch.zeros(images.size(0), 2 * TrainOptions.GRU_LAYERS, TrainOptions.HIDDEN_SIZE, device="cuda")
            for i in range(0, images.size(1), TrainOptions.WINDOW_JUMP):          # test.py:96-99
                if i + TrainOptions.TRAIN_WINDOW > images.size(1):
                    break
                image_chunk = images[:, i:i + TrainOptions.TRAIN_WINDOW]
                label_base_chunk = label_base[:, i:i + TrainOptions.TRAIN_WINDOW]
                label_rle_chunk = label_rle[:, i:i + TrainOptions.TRAIN_WINDOW]
                output_base, output_rle, hidden = transducer_model(image_chunk, hidden)
                loss_base = criterion_base(output_base.contiguous().view(-1, num_base_classes), label_base_chunk.contiguous().view(-1))
                loss_rle = criterion_rle(output_rle.contiguous().view(-1, num_rle_classes), label_rle_chunk.contiguous().view(-1))
                base_cm += _confusion(output_base, label_base_chunk, num_base_classes)
                rle_cm += _confusion(output_rle, label_rle_chunk, num_rle_classes)
                total_loss += (loss_base + loss_rle).item()
                total_loss_rle += loss_rle.item()
                total_images += images.size(0)
    base_cm, rle_cm = base_cm.cpu().numpy(), rle_cm.cpu().numpy()
    base_accuracy = 100.0 * np.trace(base_cm) / max(1.0, base_cm.sum())
    rle_accuracy = 100.0 * np.trace(rle_cm) / max(1.0, rle_cm.sum())
    avg_loss = total_loss / total_images if total_images else 0
    sys.stderr.write(TextColor.YELLOW + '\nTest Loss: ' + str(avg_loss) + " Base acc: " + str(round(base_accuracy, 4))
                     + " RLE acc: " + str(round(rle_accuracy, 4)) + "\n" + TextColor.END)
    # (the reference returns accuracy = 0: it never updates the variable, test.py:72,165; the per-head accuracies are extra keys)
    return {'loss': avg_loss, 'accuracy': 0, 'base_confusion_matrix': base_cm, 'rle_confusion_matrix': rle_cm,
            'base_accuracy': base_accuracy, 'rle_accuracy': rle_accuracy}
