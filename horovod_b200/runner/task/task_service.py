"""Task service started on each host for the interface probe (reference horovod/runner/task/task_service.py)."""
from horovod_b200.runner.common.service import task_service


class HorovodRunTaskService(task_service.BasicTaskService):
    NAME_FORMAT = 'horovod task service #%d'

    def __init__(self, index, key, nics):
        super(HorovodRunTaskService, self).__init__(HorovodRunTaskService.NAME_FORMAT % index, index, key, nics)


class HorovodRunTaskClient(task_service.BasicTaskClient):
    def __init__(self, index, task_addresses, key, verbose, match_intf=False, attempts=3):
        super(HorovodRunTaskClient, self).__init__(HorovodRunTaskService.NAME_FORMAT % index, task_addresses, key, verbose,
                                                   match_intf=match_intf, attempts=attempts)
        self.index = index
