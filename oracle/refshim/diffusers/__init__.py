"""Test-only stand-in for the `diffusers==0.24.0.dev0` symbols that
/root/reference/models/{controlnet,unet_2d_blocks}.py import.

TEST INFRASTRUCTURE ONLY.  diffusers is not installable in this image (no
network), so this package restates the leaf-module semantics the SD-1.x
configuration exercises so that the reference's own model
wiring can be executed UNMODIFIED on CPU to generate golden fixtures
(oracle/make_golden.py).  Nothing in the product path imports it.
"""


class DDPMScheduler:  # only imported for a type hint in the reference
    pass
