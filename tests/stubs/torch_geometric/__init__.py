"""TEST-ONLY stand-in for torch_geometric (not installed in this image, no network).

Only used by tools/make_golden.py in the build container to import the UNMODIFIED
reference modules from /root/reference and dump golden vectors.  Never imported by
the product package.  Semantics follow PyG 2.0.x as documented in SURVEY.md 8(c).
"""
