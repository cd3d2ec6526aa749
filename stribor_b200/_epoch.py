"""Structure epoch: a process-wide counter that moves whenever something a cached call plan depends on may have
changed without a parameter's ``(data_ptr, _version)`` showing it -- a parameter / buffer / sub-module registered on
ANY module (torch's global registration hooks), an attribute set on one of this package's modules, a packed weight
image dropped (``invalidate_packed``, ``.to()``, ``train()``, ``load_state_dict``).  ``flow.run_chain`` keeps the
ctypes layer array of a flow for as long as the epoch and the parameters' pointers / versions stand still."""
import torch.nn.modules.module as _m

value = 0


def bump(*_args, **_kwargs):
    global value
    value += 1
    return None          # registration hooks: "leave the registered object unchanged"


_m.register_module_parameter_registration_hook(bump)
_m.register_module_module_registration_hook(bump)
_m.register_module_buffer_registration_hook(bump)


class Tracked:
    """Mixin (before ``nn.Module`` in the MRO): any attribute assignment moves the epoch."""

    def __setattr__(self, name, value_):
        bump()
        super().__setattr__(name, value_)
