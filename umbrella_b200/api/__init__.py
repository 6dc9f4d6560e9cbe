"""Request front of the engines: the reference's socket protocol (umbrella/api/) and, above it, a front that spreads
independent requests over one engine per GPU (replicas only, no data-path collective)."""
