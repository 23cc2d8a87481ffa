"""Name-compatible entry point for the reference's ``initial_state.py`` (train_initial_state, simple_collate,
speaker_state_dict, parse_speaker_state, filter_unk, filter_except); the implementation lives in :mod:`lina_speech_b200.tuning`."""
from .tuning import (filter_except, filter_unk, parse_speaker_state, simple_collate, speaker_state_dict,
                     train_initial_state)

__all__ = ["train_initial_state", "simple_collate", "speaker_state_dict", "parse_speaker_state", "filter_unk", "filter_except"]
