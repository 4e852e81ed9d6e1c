"""tf.nn.* used on the path."""
import torch

relu = torch.relu
softplus = torch.nn.functional.softplus
sigmoid = torch.sigmoid


def compute_average_loss(per_example_loss, sample_weight=None, global_batch_size=None):
    """sum(per_example_loss) / global_batch_size."""
    if sample_weight is not None:
        per_example_loss = per_example_loss * sample_weight
    n = per_example_loss.shape[0] if global_batch_size is None else global_batch_size
    return per_example_loss.sum() / n
