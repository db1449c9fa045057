class PlacementGroupSchedulingStrategy:
    def __init__(self, placement_group, placement_group_bundle_index=-1):
        self.placement_group, self.placement_group_bundle_index = placement_group, placement_group_bundle_index
