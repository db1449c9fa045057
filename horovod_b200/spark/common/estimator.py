"""Shared estimator / model skeleton.

Role parity: horovod/spark/common/estimator.py (`HorovodEstimator.fit` / `fit_on_parquet` / `_fit_on_prepared_data` /
`_has_checkpoint` :26-95, `HorovodModel.transform` :97-112).  The reference derives from pyspark.ml `Estimator` / `Model`;
these classes only need a DataFrame-like input (Spark or pandas) and a Backend, and become Spark ML pipeline stages through
duck typing (`fit`, `transform`, `copy`).
"""
import time
import uuid

import numpy as np

from horovod_b200.spark.common import util
from horovod_b200.spark.common.backend import LocalBackend, SparkBackend
from horovod_b200.spark.common.params import EstimatorParams, ModelParams
from horovod_b200.spark.common.store import Store


class HorovodEstimator(EstimatorParams):
    """fit(df) -> HorovodModel.  Framework subclasses implement `_fit_on_prepared_data` and `_check_framework_params`."""

    REQUIRED = ('model', 'feature_cols', 'label_cols', 'store')

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._normalise()

    def _normalise(self):
        store = self._get('store')
        if isinstance(store, str):
            self._values['store'] = Store.create(store)

    def setParams(self, **kwargs):
        super().setParams(**kwargs)
        if '_values' in self.__dict__ and 'store' in kwargs:
            self._normalise()
        return self

    # -- validation -------------------------------------------------------------------------------------------------------------
    def _check_params(self):
        missing = [name for name in self.REQUIRED if not self._get(name)]
        if missing:
            raise ValueError('%s: required parameter(s) missing: %s' % (type(self).__name__, ', '.join(missing)))
        if self._get('backend') is not None and self._get('num_proc') is not None:
            raise ValueError('At most one of parameters "backend" and "num_proc" may be specified')
        util.check_validation(self._get('validation'))
        self._check_framework_params()

    def _check_framework_params(self):
        pass

    def _get_or_create_backend(self, df=None):
        backend = self._get('backend')
        if backend is not None:
            return backend
        if df is not None and util.is_spark_df(df):
            return SparkBackend(self._get('num_proc'))
        return LocalBackend(self._get('num_proc') or 1)

    def _new_run_id(self):
        return self._get('run_id') or 'run_' + time.strftime('%Y%m%d_%H%M%S') + '_' + uuid.uuid4().hex[:6]

    def _has_checkpoint(self, run_id):
        store = self._get('store')
        path = store.get_checkpoint_path(run_id)
        return path is not None and store.exists(path)

    def _read_checkpoint(self, run_id):
        store = self._get('store')
        return store.read(store.get_checkpoint_path(run_id)) if self._has_checkpoint(run_id) else None

    # -- training entry points ------------------------------------------------------------------------------------------------
    def fit(self, df, params=None):
        """Materialises `df` in the store, trains on `backend.num_processes()` ranks, returns the model transformer."""
        est = self.copy(params) if params else self
        est._check_params()
        backend = est._get_or_create_backend(df)
        with util.prepare_data(backend.num_processes(), est._get('store'), df,
                               label_columns=est._get('label_cols'), feature_columns=est._get('feature_cols'),
                               validation=est._get('validation'), sample_weight_col=est._get('sample_weight_col'),
                               partitions_per_process=est._get('partitions_per_process'), random_seed=est._get('random_seed') or 0,
                               verbose=est._get('verbose'), keep=False) as dataset:
            return est._fit_checked(backend, dataset)

    def fit_on_parquet(self, params=None, dataset_idx=None):
        """Trains on Parquet that is already at the store's train (and validation) data path."""
        est = self.copy(params) if params else self
        est._check_params()
        backend = est._get_or_create_backend()
        return est._fit_checked(backend, util.existing_dataset(est._get('store'), dataset_idx))

    def _fit_checked(self, backend, dataset):
        util.check_shape_compatibility(dataset.metadata, self._get('feature_cols'), self._get('label_cols'),
                                       input_shapes=self._get('input_shapes'), label_shapes=self._get('label_shapes'))
        return self._fit_on_prepared_data(backend, dataset)

    def _fit_on_prepared_data(self, backend, dataset):
        raise NotImplementedError

    def _row_shapes(self):
        shapes = {}
        for cols, declared in ((self._get('feature_cols'), self._get('input_shapes')), (self._get('label_cols'), self._get('label_shapes'))):
            if declared:
                shapes.update({c: list(s) for c, s in zip(cols, declared)})
        return shapes


class HorovodModel(ModelParams):
    """Transformer: `transform(df)` appends one prediction column per label column.  Subclasses implement
    `_predict(columns: dict name -> numpy array) -> list of numpy arrays` (one per output column)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if not self._get('output_cols') and self._get('label_columns'):
            self._values['output_cols'] = [c + '__output' for c in self._get('label_columns')]

    def getHistory(self):
        return self._get('history') or []

    def getModel(self):
        return self._get('model')

    def _predict(self, columns):
        raise NotImplementedError

    def _predict_pandas(self, pdf):
        feats = self._get('feature_columns')
        util.check_columns(pdf, feats, 'Feature')
        pdf = pdf.copy()
        outputs = {name: [] for name in self._get('output_cols')}
        step = self._get('batch_size')
        for start in range(0, len(pdf), step):
            chunk = pdf.iloc[start:start + step]
            cols = {c: np.stack([np.asarray(util._densify(v)) for v in chunk[c]]) for c in feats}
            for name, pred in zip(self._get('output_cols'), self._predict(cols)):
                pred = np.asarray(pred)
                if pred.ndim > 1 and pred.shape[-1] == 1:
                    pred = pred.reshape(pred.shape[:-1])
                outputs[name].extend(pred.tolist())
        for name, values in outputs.items():
            pdf[name] = values
        return pdf

    def transform(self, df, params=None):
        model = self.copy(params) if params else self
        if not util.is_spark_df(df):
            return model._predict_pandas(df)
        sample = model._predict_pandas(df.limit(1).toPandas())
        schema = df.sparkSession.createDataFrame(sample).schema

        def per_partition(frames):
            for pdf in frames:
                yield model._predict_pandas(pdf)
        return df.mapInPandas(per_partition, schema=schema)
