"""Which DataFrames are already materialised in which store (reference horovod/spark/common/cache.py `TrainingDataCache`
:21-90, behind `util.prepare_data` and `util.clear_training_cache`)."""
import threading
import weakref


class TrainingDataCache:
    """Remembers which DataFrame objects are already materialised in which store so that fitting several models (a
    hyper-parameter search) on the same DataFrame writes the Parquet once.  Entries die with their DataFrame."""

    def __init__(self):
        self._lock = threading.Lock()
        self._entries = {}       # key -> (weakref to df or None, PreparedDataset, users)

    @staticmethod
    def _key(df, store, validation, columns, num_files):
        return (id(df), getattr(store, 'prefix_path', id(store)), repr(validation), tuple(columns), num_files)

    def lookup(self, df, store, validation, columns, num_files):
        key = self._key(df, store, validation, columns, num_files)
        with self._lock:
            hit = self._entries.get(key)
            if hit is None:
                return key, None
            ref, dataset, users = hit
            if ref is not None and ref() is not df:          # the id was recycled by another object
                del self._entries[key]
                return key, None
            if not store.exists(dataset.train_path):
                del self._entries[key]
                return key, None
            self._entries[key] = (ref, dataset, users + 1)
            return key, dataset

    def insert(self, key, df, dataset):
        try:
            ref = weakref.ref(df)
        except TypeError:
            ref = None
        with self._lock:
            self._entries[key] = (ref, dataset, 1)

    def release(self, key):
        with self._lock:
            hit = self._entries.get(key)
            if hit:
                self._entries[key] = (hit[0], hit[1], max(0, hit[2] - 1))

    def clear(self, store=None):
        """Forgets (and deletes from `store`) every materialised dataset nobody is training on."""
        with self._lock:
            for key, (ref, dataset, users) in list(self._entries.items()):
                if users == 0:
                    if store is not None:
                        store.delete(dataset.train_path)
                        if dataset.val_path:
                            store.delete(dataset.val_path)
                    del self._entries[key]
