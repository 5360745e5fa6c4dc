"""Model container and drivers (reference denet/model)."""
