import concurrent.futures

import ray


class PlacementGroup:
    def __init__(self, bundles, strategy):
        self.bundle_specs, self.strategy, self.removed = bundles, strategy, False

    def ready(self):
        fut = concurrent.futures.Future()
        need = {}
        for b in self.bundle_specs:
            for k, v in b.items():
                need[k] = need.get(k, 0) + v
        if all(ray._state['cluster'].get(k, 0) >= v for k, v in need.items()):
            fut.set_result(self)
        return ray.ObjectRef(fut)          # never completes when the cluster is too small


def placement_group(bundles, strategy='PACK'):
    pg = PlacementGroup(bundles, strategy)
    ray._state['groups'].append(pg)
    return pg


def get_current_placement_group():
    return ray._state['current_pg']


def remove_placement_group(pg):
    pg.removed = True
