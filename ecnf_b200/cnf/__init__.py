"""Reference-shaped API of ecnf/cnf (same names and argument order), batched, backed by libecnf_b200.so."""
from .core import FlowMatchingCNF, optimal_transport_conditional_vf  # noqa: F401
from .build_cnf import build_cnf, get_timestep_embedding  # noqa: F401
from .sample_and_log_prob import sample_cnf, get_log_prob, sample_and_log_prob_cnf  # noqa: F401
from .loss import flow_matching_loss_fn, flow_matching_loss_and_grad_fn  # noqa: F401
from .gradient_step import TrainingState, flow_matching_update_fn  # noqa: F401
