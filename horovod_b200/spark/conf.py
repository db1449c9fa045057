"""Spark properties that matter for elastic jobs (reference horovod/spark/conf.py).

An elastic job survives task failures by re-planning itself, so Spark must neither give up after a few failed task attempts
nor exclude executors / nodes on its own.  Each constant is a (property, value) pair for `SparkConf.set` or
`spark-submit --conf property=value`; `elastic_conf()` returns the combination `run_elastic` expects.
"""
SPARK_CONF_MAX_INT = '2147483647'
SPARK_CONF_MAX_INT_MINUS_ONE = '2147483646'

# the job has its own retry limit (reset_limit): Spark retries every failed task
SPARK_CONF_ALWAYS_RESTART_FAILED_TASK = ('spark.task.maxFailures', SPARK_CONF_MAX_INT)

# Spark's executor / node exclusion ("blacklist")
SPARK_CONF_BLACKLIST_DISABLED = ('spark.blacklist.enabled', 'false')
SPARK_CONF_BLACKLIST_ENABLED = ('spark.blacklist.enabled', 'true')

# within a stage: may an executor on which a task failed run other tasks / may the other executors of its node be used
SPARK_CONF_REUSE_FAILED_EXECUTOR = ('spark.blacklist.stage.maxFailedTasksPerExecutor', SPARK_CONF_MAX_INT)
SPARK_CONF_DONT_REUSE_FAILED_EXECUTOR = ('spark.blacklist.stage.maxFailedTasksPerExecutor', '1')
SPARK_CONF_REUSE_FAILING_NODE = ('spark.blacklist.stage.maxFailedExecutorsPerNode', SPARK_CONF_MAX_INT_MINUS_ONE)
SPARK_CONF_DONT_REUSE_FAILING_NODE = ('spark.blacklist.stage.maxFailedExecutorsPerNode', '1')

# for one task: how often it may be retried on the same executor / node
SPARK_CONF_REUSE_EXECUTOR_ALWAYS_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerExecutor', SPARK_CONF_MAX_INT)
SPARK_CONF_REUSE_EXECUTOR_ONCE_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerExecutor', '2')
SPARK_CONF_DONT_REUSE_EXECUTOR_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerExecutor', '1')
SPARK_CONF_REUSE_NODE_ALWAYS_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerNode', SPARK_CONF_MAX_INT_MINUS_ONE)
SPARK_CONF_REUSE_NODE_ONCE_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerNode', '2')
SPARK_CONF_DONT_REUSE_NODE_FOR_SAME_TASK = ('spark.blacklist.task.maxTaskAttemptsPerNode', '1')

# across the application (with dynamic allocation only application-wide exclusions give executors back to the cluster manager)
SPARK_CONF_REUSE_FAILED_EXECUTOR_IN_APP = ('spark.blacklist.application.maxFailedTasksPerExecutor', SPARK_CONF_MAX_INT)
SPARK_CONF_DONT_REUSE_FAILED_EXECUTOR_IN_APP = ('spark.blacklist.application.maxFailedTasksPerExecutor', '1')
SPARK_CONF_REUSE_FAILING_NODE_IN_APP = ('spark.blacklist.application.maxFailedExecutorsPerNode', SPARK_CONF_MAX_INT)
SPARK_CONF_DONT_REUSE_FAILING_NODE_IN_APP = ('spark.blacklist.application.maxFailedExecutorsPerNode', '1')

# Spark's own defaults for the properties above
SPARK_CONF_DEFAULT_VALUES = {
    'spark.task.maxFailures': '4',
    'spark.blacklist.enabled': 'false',
    'spark.blacklist.stage.maxFailedTasksPerExecutor': '2',
    'spark.blacklist.stage.maxFailedExecutorsPerNode': '2',
    'spark.blacklist.task.maxTaskAttemptsPerExecutor': '1',
    'spark.blacklist.task.maxTaskAttemptsPerNode': '2',
    'spark.blacklist.application.maxFailedTasksPerExecutor': '2',
    'spark.blacklist.application.maxFailedExecutorsPerNode': '2',
}


def elastic_conf(reuse_failed_executors=True):
    """Properties for a SparkSession that runs `horovod_b200.spark.run_elastic`: unlimited task retries and either no
    exclusion at all, or exclusion that never triggers."""
    pairs = [SPARK_CONF_ALWAYS_RESTART_FAILED_TASK]
    if reuse_failed_executors:
        pairs.append(SPARK_CONF_BLACKLIST_DISABLED)
    else:
        pairs += [SPARK_CONF_BLACKLIST_ENABLED, SPARK_CONF_DONT_REUSE_FAILED_EXECUTOR, SPARK_CONF_DONT_REUSE_FAILING_NODE,
                  SPARK_CONF_DONT_REUSE_EXECUTOR_FOR_SAME_TASK]
    return dict(pairs)


def check_elastic_conf(conf_get, warn=None):
    """`conf_get(property, default)` (e.g. `spark.sparkContext.getConf().get`): returns the properties whose current value
    would make Spark abort an elastic job on its own, i.e. fewer task retries than the job's reset limit can need."""
    problems = {}
    value = conf_get('spark.task.maxFailures', SPARK_CONF_DEFAULT_VALUES['spark.task.maxFailures'])
    if int(value) < int(SPARK_CONF_MAX_INT):
        problems['spark.task.maxFailures'] = value
    if problems and warn is not None:
        warn('Spark gives up before the elastic job does: set %s' % ', '.join('%s=%s' % kv for kv in elastic_conf().items()))
    return problems
