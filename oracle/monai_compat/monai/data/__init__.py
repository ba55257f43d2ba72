"""monai.data subset: decollate_batch for tensors."""


def decollate_batch(batch):
    return [batch[i] for i in range(batch.shape[0])]
