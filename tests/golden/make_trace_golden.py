#!/usr/bin/env python3
"""Write tests/golden/saved_model_trace.pt: the reference's own TorchScript export of the shipped prod_net, made exactly the way
training_scripts/make_torchscript_model.py:17-34 makes it (load_and_glue_nets -> eval -> torch.jit.trace on randn [1,3,144,256]
-> save), by the UNMODIFIED reference imported from /root/reference.  Build container only.

    python tests/golden/make_trace_golden.py
"""
import os
import sys

REF = "/root/reference"
sys.path.insert(0, REF)
import torch  # noqa: E402
import frameID.net as ref_net  # noqa: E402

assert ref_net.__file__.startswith(REF), ref_net.__file__
torch.manual_seed(0)
net, params = ref_net.load_default_net()
net.eval()
traced = torch.jit.trace(net, torch.randn([1, 3, 144, 256]))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "saved_model_trace.pt")
traced.save(out)
print(out, os.path.getsize(out), "bytes;", len(traced.state_dict()), "tensors;", params)
