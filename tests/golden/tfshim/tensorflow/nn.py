"""tf.nn.* used on the path."""
import torch

relu = torch.relu
softplus = torch.nn.functional.softplus
sigmoid = torch.sigmoid
