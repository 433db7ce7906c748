"""Host-side pattern description (graph, measurements, flow, templates)."""
from . import cluster_templates as templates
from .causal_flow import Flow, check_if_flow, find_cflow
from .circuit import MBQCircuit, hstack, merge, vstack
from .graph import GraphState
from .measurement import ControlledMent, ControlMent, Measurement, Ment, MentOutcome

__all__ = ["GraphState", "MBQCircuit", "Ment", "Measurement", "ControlMent", "ControlledMent", "MentOutcome", "Flow", "find_cflow", "check_if_flow",
           "hstack", "vstack", "merge", "templates"]
