from .multilevel_region_proposal_network import map_rois_to_fpn_levels  # noqa: F401
