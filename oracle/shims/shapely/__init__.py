"""Stand-in for shapely (only imported, never called on the roadmap path)."""
