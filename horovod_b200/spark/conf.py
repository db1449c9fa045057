"""Spark properties that matter for elastic jobs (role parity: horovod/spark/conf.py).

An elastic job survives task failures by re-planning itself, so Spark must neither give up after a few failed task attempts nor
exclude executors / nodes on its own.  Every public constant is a `(property, value)` pair for `SparkConf.set` or
`spark-submit --conf property=value`.  They are generated from one table (`_EXCLUSION_KNOBS`): per Spark exclusion ("blacklist")
property the table lists the named choices the reference exposes — e.g. `SPARK_CONF_REUSE_FAILED_EXECUTOR` /
`SPARK_CONF_DONT_REUSE_FAILED_EXECUTOR` are the two rows of `stage.maxFailedTasksPerExecutor`.  `elastic_conf()` returns the
combination `run_elastic` expects, `check_elastic_conf()` tells which current settings would make Spark abort the job.
"""
SPARK_CONF_MAX_INT = str(2 ** 31 - 1)
SPARK_CONF_MAX_INT_MINUS_ONE = str(2 ** 31 - 2)

_PREFIX = 'spark.blacklist.'
# property (below spark.blacklist.) -> Spark's default, then {constant suffix: value}
#   stage.*        within a stage: may an executor on which a task failed run other tasks / may its node's other executors be used
#   task.*         for one task: how often it may be retried on the same executor / node
#   application.*  across the application (with dynamic allocation only these give executors back to the cluster manager)
_EXCLUSION_KNOBS = {
    'enabled': ('false', {'BLACKLIST_DISABLED': 'false', 'BLACKLIST_ENABLED': 'true'}),
    'stage.maxFailedTasksPerExecutor': ('2', {'REUSE_FAILED_EXECUTOR': SPARK_CONF_MAX_INT, 'DONT_REUSE_FAILED_EXECUTOR': '1'}),
    'stage.maxFailedExecutorsPerNode': ('2', {'REUSE_FAILING_NODE': SPARK_CONF_MAX_INT_MINUS_ONE, 'DONT_REUSE_FAILING_NODE': '1'}),
    'task.maxTaskAttemptsPerExecutor': ('1', {'REUSE_EXECUTOR_ALWAYS_FOR_SAME_TASK': SPARK_CONF_MAX_INT,
                                              'REUSE_EXECUTOR_ONCE_FOR_SAME_TASK': '2', 'DONT_REUSE_EXECUTOR_FOR_SAME_TASK': '1'}),
    'task.maxTaskAttemptsPerNode': ('2', {'REUSE_NODE_ALWAYS_FOR_SAME_TASK': SPARK_CONF_MAX_INT_MINUS_ONE,
                                          'REUSE_NODE_ONCE_FOR_SAME_TASK': '2', 'DONT_REUSE_NODE_FOR_SAME_TASK': '1'}),
    'application.maxFailedTasksPerExecutor': ('2', {'REUSE_FAILED_EXECUTOR_IN_APP': SPARK_CONF_MAX_INT,
                                                    'DONT_REUSE_FAILED_EXECUTOR_IN_APP': '1'}),
    'application.maxFailedExecutorsPerNode': ('2', {'REUSE_FAILING_NODE_IN_APP': SPARK_CONF_MAX_INT,
                                                    'DONT_REUSE_FAILING_NODE_IN_APP': '1'}),
}

# the job has its own retry limit (reset_limit): Spark retries every failed task
SPARK_CONF_ALWAYS_RESTART_FAILED_TASK = ('spark.task.maxFailures', SPARK_CONF_MAX_INT)
SPARK_CONF_DEFAULT_VALUES = {'spark.task.maxFailures': '4'}
for _prop, (_default, _choices) in _EXCLUSION_KNOBS.items():
    SPARK_CONF_DEFAULT_VALUES[_PREFIX + _prop] = _default
    for _suffix, _value in _choices.items():
        globals()['SPARK_CONF_' + _suffix] = (_PREFIX + _prop, _value)
del _prop, _default, _choices, _suffix, _value


def elastic_conf(reuse_failed_executors=True):
    """Properties for a SparkSession that runs `horovod_b200.spark.run_elastic`: unlimited task retries and either no exclusion
    at all, or exclusion of exactly the executor / node a task failed on."""
    g = globals()
    names = ['ALWAYS_RESTART_FAILED_TASK']
    names += ['BLACKLIST_DISABLED'] if reuse_failed_executors else \
        ['BLACKLIST_ENABLED', 'DONT_REUSE_FAILED_EXECUTOR', 'DONT_REUSE_FAILING_NODE', 'DONT_REUSE_EXECUTOR_FOR_SAME_TASK']
    return dict(g['SPARK_CONF_' + n] for n in names)


def check_elastic_conf(conf_get, warn=None):
    """`conf_get(property, default)` (e.g. `spark.sparkContext.getConf().get`): returns the properties whose current value would
    make Spark abort an elastic job on its own, i.e. fewer task retries than the job's reset limit can need."""
    prop = SPARK_CONF_ALWAYS_RESTART_FAILED_TASK[0]
    value = conf_get(prop, SPARK_CONF_DEFAULT_VALUES[prop])
    problems = {prop: value} if int(value) < int(SPARK_CONF_MAX_INT) else {}
    if problems and warn is not None:
        warn('Spark gives up before the elastic job does: set %s' % ', '.join('%s=%s' % kv for kv in elastic_conf().items()))
    return problems
