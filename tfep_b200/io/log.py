"""Per-sample training / evaluation log in the reference's on-disk layout (reference tfep/io/log.py:34-643), so that
analysis scripts written for tfep read what this package produces (potentials, log_det_J, sample indices ->
``fep_estimator`` / ``bootstrap``) and vice versa.

Layout (version '0.1'):

* ``<dir>/metadata.json``: ``{"batch_size", "n_samples_per_epoch", "version"}``;
* ``<dir>/train/epoch-E.npz``: one 1-D array of length ``n_samples_per_epoch`` per logged quantity, entry
  ``batch_idx * batch_size + i`` = sample ``i`` of batch ``batch_idx``, plus the boolean array ``__mask`` marking the
  entries written so far;
* ``<dir>/eval/step-S.npz``: one 1-D array per quantity for the model after ``S`` optimisation steps, batches
  appended in the order they were saved (or merged on the sample indices with ``update=True``).

Tensors may live on any device; they are detached and brought to the host when saved.  Not thread / process safe
(like the reference): one writer per directory.
"""

import json
import os
import warnings

import numpy as np
import torch


class _Archive:
    """The ``.npz`` file currently held in memory for one kind of data ('train' or 'eval')."""

    def __init__(self, directory, prefix):
        self.directory, self.prefix = directory, prefix
        self.index, self.arrays = None, None

    def path(self):
        return os.path.join(self.directory, f'{self.prefix}-{self.index}.npz')

    def open(self, index, empty):
        """Make ``index`` the current archive: read it from disk if it exists, else start from ``empty()``."""
        if self.index == index:
            return
        self.index = index
        if os.path.isfile(self.path()):
            with np.load(self.path()) as f:
                self.arrays = {k: f[k] for k in f.files}
        else:
            self.arrays = empty()

    def flush(self):
        np.savez_compressed(self.path(), **self.arrays)


class TFEPLogger:
    """Store and retrieve per-sample quantities (potential energies, log_det_J, CVs, sample indices) by epoch, batch or
    optimisation step.  Same constructor, methods, file names and array layout as the reference's ``TFEPLogger``.

    Parameters
    ----------
    save_dir_path : str
    data_loader : torch.utils.data.DataLoader, optional
        Needed when the directory holds no ``metadata.json`` yet: gives ``batch_size``, ``drop_last`` and the dataset
        length.  Ignored when resuming.
    train_subdir_name, eval_subdir_name : str
    """

    VERSION = '0.1'
    METADATA_FILE_NAME = 'metadata.json'
    INDEX_NAMES = ['trajectory_sample_index', 'dataset_sample_index']
    MASK_NAME = '__mask'

    def __init__(self, save_dir_path='tfep_logs', data_loader=None, train_subdir_name='train', eval_subdir_name='eval'):
        self._save_dir_path = os.path.realpath(save_dir_path)
        self._train = _Archive(os.path.join(save_dir_path, train_subdir_name), 'epoch')
        self._eval = _Archive(os.path.join(save_dir_path, eval_subdir_name), 'step')
        metadata_path = os.path.join(save_dir_path, self.METADATA_FILE_NAME)
        resume = os.path.isfile(metadata_path)
        if resume:
            with open(metadata_path, 'r') as f:
                metadata = json.load(f)
            self._batch_size = metadata['batch_size']
            self._n_samples_per_epoch = metadata['n_samples_per_epoch']
        elif data_loader is None:
            raise ValueError("When creating a new logger, 'data_loader' must be passed.")
        else:
            self._batch_size, self._n_samples_per_epoch = self._epoch_shape(data_loader)
        for d in (save_dir_path, self._train.directory, self._eval.directory):
            os.makedirs(d, exist_ok=True)
        if not resume:
            with open(metadata_path, 'w') as f:
                json.dump({'batch_size': self.batch_size, 'n_samples_per_epoch': self.n_samples_per_epoch,
                           'version': self.VERSION}, f)

    @staticmethod
    def _epoch_shape(data_loader):
        batch_size = data_loader.batch_size
        if batch_size is None:                       # a batch sampler was passed to the loader
            batch_size = data_loader.batch_sampler.batch_size
            drop_last = data_loader.batch_sampler.drop_last
        else:
            drop_last = data_loader.drop_last
        n = len(data_loader.dataset)
        return batch_size, (n - n % batch_size) if drop_last else n

    # -- geometry of an epoch ----------------------------------------------------------------------------------
    @property
    def batch_size(self):
        return self._batch_size

    @property
    def n_samples_per_epoch(self):
        return self._n_samples_per_epoch

    @property
    def n_batches_per_epoch(self):
        return int(np.ceil(self.n_samples_per_epoch / self.batch_size))

    @property
    def save_dir_path(self):
        return self._save_dir_path

    def _resolve(self, step_idx, epoch_idx, batch_idx, need_batch):
        """(step, epoch, batch) from either a step index or an epoch (+ batch) index."""
        per_epoch = self.n_batches_per_epoch
        if step_idx is not None:
            epoch_idx, batch_idx = divmod(step_idx, per_epoch)
        elif epoch_idx is None:
            raise ValueError("Either step_idx or epoch_idx must be passed.")
        elif batch_idx is None:
            if need_batch:
                raise ValueError("To save tensors either 'step_idx' or both "
                                 "'epoch_idx' and 'batch_idx' must be passed.")
        else:
            step_idx = epoch_idx * per_epoch + batch_idx
        return step_idx, epoch_idx, batch_idx

    def _open_train(self, epoch_idx):
        self._train.open(epoch_idx, lambda: {self.MASK_NAME: np.full(self.n_samples_per_epoch, fill_value=False)})

    @classmethod
    def _warn_if_no_indices(cls, tensors):
        if not any(name in tensors for name in cls.INDEX_NAMES):
            warnings.warn(("tensors does not contain any sample indices among the "
                           "following attributes: {}. Without it, it might be "
                           "difficult to match training and evaluation configurations "
                           "to their reference potential.").format(cls.INDEX_NAMES))

    @staticmethod
    def _host(value):
        return value.detach().cpu().numpy()

    # -- writing ----------------------------------------------------------------------------------------------------
    def save_train_tensors(self, tensors, step_idx=None, epoch_idx=None, batch_idx=None):
        """Save the per-sample tensors of one training batch (or of a whole epoch if no batch index resolves)."""
        self._warn_if_no_indices(tensors)
        _, epoch_idx, batch_idx = self._resolve(step_idx, epoch_idx, batch_idx, need_batch=False)
        self._open_train(epoch_idx)
        data = self._train.arrays
        mask = data[self.MASK_NAME]
        for name, value in tensors.items():
            value = self._host(value)
            if batch_idx is None:
                data[name] = value
                mask[:] = True
            else:
                if name not in data:
                    data[name] = np.empty(self.n_samples_per_epoch, dtype=value.dtype)
                first = self.batch_size * batch_idx
                data[name][first:first + len(value)] = value
                mask[first:first + len(value)] = True
        self._train.flush()

    def save_eval_tensors(self, tensors, step_idx=None, epoch_idx=None, batch_idx=None, update=False):
        """Append the per-sample tensors of one evaluation batch to the archive of the step; with ``update`` entries
        whose sample index is already stored are overwritten instead of appended."""
        self._warn_if_no_indices(tensors)
        step_idx, _, _ = self._resolve(step_idx, epoch_idx, batch_idx, need_batch=True)
        self._eval.open(step_idx, dict)
        data = self._eval.arrays
        names = list(tensors.keys()) if len(data) == 0 else list(data.keys())
        try:
            new = {n: self._host(tensors[n]) for n in names}
        except KeyError:
            raise KeyError("'tensors' must include all the following Tensors: " + str(names))
        if update:
            for index_name in self.INDEX_NAMES:
                if index_name not in new:
                    continue
                _, in_new, in_old = np.intersect1d(new[index_name], data[index_name], assume_unique=True,
                                                   return_indices=True)
                if len(in_new) == 0:
                    break
                for n in names:
                    data[n][in_old] = new[n][in_new]
                    new[n] = np.delete(new[n], in_new)
        for n in names:
            data[n] = np.concatenate((data[n], new[n])) if n in data else new[n]
        self._eval.flush()

    # -- reading ----------------------------------------------------------------------------------------------------
    def _nan_mask(self, arrays, remove_nans, written=None):
        if remove_nans is False:
            return written
        if remove_nans is True:
            mask = None
            for name, value in arrays.items():
                if name != self.MASK_NAME:
                    mask = ~np.isnan(value) if mask is None else mask & ~np.isnan(value)
        else:
            mask = ~np.isnan(arrays[remove_nans])
        return mask if written is None else mask & written

    def read_train_tensors(self, names=None, step_idx=None, epoch_idx=None, batch_idx=None, remove_nans=False,
                           as_numpy=False):
        """The saved entries of an epoch (or of one of its batches): dict name -> 1-D tensor."""
        _, epoch_idx, batch_idx = self._resolve(step_idx, epoch_idx, batch_idx, need_batch=False)
        self._open_train(epoch_idx)
        data = self._train.arrays
        if names is None:
            names = [k for k in data if k != self.MASK_NAME]
        mask = self._nan_mask(data, remove_nans, written=data[self.MASK_NAME])
        window = slice(None) if batch_idx is None else slice(self.batch_size * batch_idx, self.batch_size * (batch_idx + 1))
        out = {n: data[n][window][mask[window]] for n in names}
        return out if as_numpy else {k: torch.tensor(v) for k, v in out.items()}

    def read_eval_tensors(self, names=None, step_idx=None, epoch_idx=None, batch_idx=None, remove_nans=False,
                          sort_by=None, as_numpy=False):
        """The saved entries of an evaluation step: dict name -> 1-D tensor.  ``sort_by`` reorders all arrays by the named
        one (e.g. 'trajectory_sample_index') and stores the new order on disk."""
        step_idx, _, _ = self._resolve(step_idx, epoch_idx, batch_idx, need_batch=True)
        self._eval.open(step_idx, dict)
        if sort_by is not None:
            order = np.argsort(self._eval.arrays[sort_by])
            self._eval.arrays = {k: v[order] for k, v in self._eval.arrays.items()}
            self._eval.flush()
        data = self._eval.arrays
        out = data if names is None else {n: data[n] for n in names}
        mask = self._nan_mask(data, remove_nans)
        if mask is not None:
            out = {k: v[mask] for k, v in out.items()}
        return out if as_numpy else {k: torch.tensor(v) for k, v in out.items()}
