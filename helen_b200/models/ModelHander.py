"""Checkpoint surface (mirrors helen/modules/python/models/ModelHander.py; the file name
keeps the reference's spelling so imports port 1:1).

.pkl = torch.save({'model_state_dict', 'model_optimizer', 'hidden_size', 'gru_layers', 'epochs'}).
"""
import os
from collections import OrderedDict

import torch

from .TransducerModel import TransducerGRU


class ModelHandler:
    @staticmethod
    def save_checkpoint(state, filename):
        torch.save(state, filename)

    @staticmethod
    def get_new_gru_model(input_channels, image_features, gru_layers, hidden_size, num_base_classes, num_rle_classes):
        return TransducerGRU(input_channels, image_features, gru_layers, hidden_size, num_base_classes,
                             num_rle_classes, bidirectional=True)

    @staticmethod
    def load_checkpoint(model_path):
        """torch.load(map_location='cpu') of a reference checkpoint (ModelHander.py:50)."""
        try:
            return torch.load(model_path, map_location='cpu')
        except Exception:
            # checkpoints that also pickle optimizer state / numpy scalars need the full unpickler
            return torch.load(model_path, map_location='cpu', weights_only=False)

    @staticmethod
    def load_simple_model(model_path, input_channels, image_features, seq_len, num_base_classes, num_rle_classes):
        """ModelHander.py:38-82: returns (model, hidden_size, gru_layers, epochs)."""
        checkpoint = ModelHandler.load_checkpoint(model_path)
        hidden_size = checkpoint['hidden_size']
        gru_layers = checkpoint['gru_layers']
        epochs = checkpoint['epochs']
        model = ModelHandler.get_new_gru_model(input_channels=input_channels, image_features=image_features,
                                               gru_layers=gru_layers, hidden_size=hidden_size,
                                               num_base_classes=num_base_classes, num_rle_classes=num_rle_classes)
        new_state = OrderedDict()
        for k, v in checkpoint['model_state_dict'].items():
            name = k[7:] if k[0:7] == 'module.' else k      # DataParallel/DDP prefix
            new_state[name] = v
        model.load_state_dict(new_state)
        model.cpu()
        return model, hidden_size, gru_layers, epochs

    @staticmethod
    def save_model(transducer_model, model_optimizer, hidden_size, layers, epoch, file_name):
        """ModelHander.py:109-133."""
        if os.path.isfile(file_name):
            os.remove(file_name)
        optimizer_state = model_optimizer.state_dict() if model_optimizer is not None else {}
        ModelHandler.save_checkpoint({
            'model_state_dict': transducer_model.state_dict(),
            'model_optimizer': optimizer_state,
            'hidden_size': hidden_size,
            'gru_layers': layers,
            'epochs': epoch,
        }, file_name)
