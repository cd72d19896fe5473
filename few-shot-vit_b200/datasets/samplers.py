"""CategoriesSampler (reference: test_phase/datasets/samplers.py:5-35): yields, per batch, the flat index tensor
[ep_per_batch * n_cls * n_per] of `ep_per_batch` episodes, each `n_cls` classes drawn without replacement and `n_per` images
per class drawn without replacement, in class-major order.  Draws come from numpy's global RNG in the reference's order
(classes first, then per class), so the same np.random.seed gives the same batches."""
import numpy as np
import torch


class CategoriesSampler:
    def __init__(self, label, n_batch, n_cls, n_per, ep_per_batch=1):
        self.n_batch, self.n_cls, self.n_per, self.ep_per_batch = n_batch, n_cls, n_per, ep_per_batch
        label = np.array(label)
        self.catlocs = [np.argwhere(label == c).reshape(-1) for c in range(max(label) + 1)]

    def __len__(self):
        return self.n_batch

    def __iter__(self):
        for _ in range(self.n_batch):
            batch = []
            for _ in range(self.ep_per_batch):
                classes = np.random.choice(len(self.catlocs), self.n_cls, replace=False)
                episode = [torch.from_numpy(np.random.choice(self.catlocs[c], self.n_per, replace=False)) for c in classes]
                batch.append(torch.stack(episode))
            yield torch.stack(batch).view(-1)
