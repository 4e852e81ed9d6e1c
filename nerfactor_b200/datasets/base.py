"""Mirror of nerfactor/datasets/base.py:25-114: file globbing, per-view loading, optional caching,
shuffling and prefetching.  `build_pipeline` returns a re-iterable `DataPipe` (one pass = one
epoch) instead of a tf.data graph: a loader thread reads / decodes the next views while the GPU
works on the current one and stages float arrays in page-locked memory, so the consumer's
host->device copies are asynchronous."""
import queue
import threading

import numpy as np


class Dataset:
    def __init__(self, config, mode, debug=False, shuffle_buffer_size=64,
                 prefetch_buffer_size=2, n_map_parallel_calls=None, seed=None):
        assert mode in ('train', 'vali', 'test'), (
            "Accepted dataset modes: 'train', 'vali', 'test', but input is %s") % mode
        self.config = config
        self.mode = mode
        self.debug = debug
        self.shuffle_buffer_size = config.getint(
            'DEFAULT', 'shuffle_buffer_size', fallback=shuffle_buffer_size)
        self.prefetch_buffer_size = prefetch_buffer_size
        self.rng = np.random.default_rng(seed)
        self.files = self._glob()
        assert self.files, "No file to process into a dataset"
        self.bs = self._get_batch_size()

    # -- to override ---------------------------------------------------------------
    def _glob(self):
        raise NotImplementedError

    def _get_batch_size(self):
        """base.py:55-67."""
        if 'bs' not in self.config['DEFAULT'].keys():
            raise ValueError(
                "Specify batch size either as 'bs' in the configuration file, "
                "or override this function to generate a value another way")
        return self.config.getint('DEFAULT', 'bs')

    def _process_example_precache(self, path):
        """Output of this function will be cached."""
        raise NotImplementedError

    def _process_example_postcache(self, *args):
        """Whatever involves randomness (ray sampling); default no-op."""
        return args

    # -- pipeline ------------------------------------------------------------------
    def build_pipeline(self, filter_predicate=None, seed=None, no_batch=False,
                       no_shuffle=False, pin_memory=True):
        """base.py:84-114.  `no_batch` is accepted for signature parity: NeRF-style datasets
        always run with no_batch=True (one view = one batch, config/*.ini), which is what a
        DataPipe element is."""
        files = sorted(self.files)
        if filter_predicate is not None:
            files = [f for f in files if filter_predicate(f)]
        cache = self.config.getboolean('DEFAULT', 'cache', fallback=False)
        shuffle = self.mode == 'train' and not no_shuffle
        rng = np.random.default_rng(seed) if seed is not None else self.rng
        return DataPipe(self, files, cache, shuffle, rng, self.prefetch_buffer_size, pin_memory)


class DataPipe:
    """One iteration = one epoch over the views.  `take(n)` keeps the first n elements (the fixed
    validation batches of trainvali.py:98-100)."""

    def __init__(self, dataset, files, cache, shuffle, rng, prefetch, pin_memory, limit=None):
        self.dataset, self.files, self.cache = dataset, files, cache
        self.shuffle, self.rng, self.prefetch = shuffle, rng, max(1, int(prefetch))
        self.pin_memory, self.limit = pin_memory, limit
        self._cached = {}

    def __len__(self):
        n = len(self.files)
        return n if self.limit is None else min(n, self.limit)

    def take(self, n):
        p = DataPipe(self.dataset, self.files, self.cache, self.shuffle, self.rng, self.prefetch,
                     self.pin_memory, limit=n)
        p._cached = self._cached
        return p

    def _load(self, path):
        if self.cache and path in self._cached:
            pre = self._cached[path]
        else:
            pre = self.dataset._process_example_precache(path)
            if self.cache:
                self._cached[path] = pre
        out = self.dataset._process_example_postcache(*pre)
        return tuple(_stage(x, self.pin_memory) for x in out)

    def __iter__(self):
        order = list(self.files)
        if self.shuffle:
            order = [order[i] for i in self.rng.permutation(len(order))]
        if self.limit is not None:
            order = order[:self.limit]
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def work():
            try:
                for path in order:
                    if stop.is_set():
                        return
                    q.put(('ok', self._load(path)))
                q.put(('end', None))
            except BaseException as e:            # surfaced in the consumer thread
                q.put(('err', e))

        t = threading.Thread(target=work, daemon=True)
        t.start()
        try:
            while True:
                kind, item = q.get()
                if kind == 'end':
                    return
                if kind == 'err':
                    raise item
                yield item
        finally:
            stop.set()
            while t.is_alive():                   # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)


def _stage(x, pin):
    """float / int arrays -> torch tensors (page-locked when a GPU is present)."""
    import torch
    if isinstance(x, np.ndarray) and x.dtype.kind in 'fiu':
        t = torch.from_numpy(np.ascontiguousarray(x))
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        return t
    return x
