"""integration/addon.cc is the N-API shim a Node deployment would load
(INTEGRATION.md).  Node is not in this image, so the addon cannot be built for
real; this keeps it from rotting: it must COMPILE (-fsyntax-only, warnings as
errors) against include/dragnet_gpu.h and a hand-written declaration stub of
the N-API calls it uses, and every dng_* function it calls must be one the
library exports."""

import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_addon_compiles_against_the_header_and_an_napi_stub():
    subprocess.check_call(
        ['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-Wextra', '-Werror',
         '-I', os.path.join(ROOT, 'tests', 'node_api_stub'),
         '-I', os.path.join(ROOT, 'include'),
         os.path.join(ROOT, 'integration', 'addon.cc')])


def test_addon_only_calls_exported_entry_points():
    src = open(os.path.join(ROOT, 'integration', 'addon.cc')).read()
    hdr = open(os.path.join(ROOT, 'include', 'dragnet_gpu.h')).read()
    called = set(re.findall(r'\b(dng_[a-z_]+)\s*\(', src))
    declared = set(re.findall(r'\b(dng_[a-z_]+)\s*\(', hdr))
    assert called and called <= declared, called - declared
    js = open(os.path.join(ROOT, 'integration', 'datasource-gpu.js')).read()
    for m in ('scanOpen', 'feedFile', 'finish'):
        assert m in js and m in src
