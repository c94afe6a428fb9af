"""Result containers with the reference's attribute layout (``src/data_containers/helper_interfaces/i_collection.py``).

The instance attributes are the on-disk schema of the reference's jsonpickle files (``_molecular_parameter``,
``_optimized_exp_values``, ``_fci_value``, ``_hf_value``, ``_label``; ``_measurements``, ``_mean``; ``data_noisy``,
``data_mitigated``, ``mu_ideal``, ``_settings``), so they keep their names; ``jsonpickle_compat`` restores instances
without calling ``__init__``.  Files written by an earlier revision of the reference carry
``_measured_eigenvalue`` / ``_measured_standard_deviation`` instead of the list of optimised values — the accessors fall
back to them.
"""
from __future__ import annotations

import statistics
from typing import Iterator, List, Tuple

import numpy as np

from .i_noise_wrapper import INoiseModel, INoiseWrapper
from .i_parameter import IParameter
from .i_wave_function import IWaveFunction

_NAN = float("NaN")


def _or_nan(value):
    return _NAN if value is None else value


class IContainer:
    """One (noise model, molecule parameter) point: every optimised energy of its restarts plus the FCI / HF references
    (reference :12-51)."""

    def __init__(self, m_param: IParameter, e_values: List[float], m_data, label: str = "<input label>"):
        self._molecular_parameter = m_param["r0"]
        self._optimized_exp_values = e_values
        self._fci_value, self._hf_value = m_data.fci_energy, m_data.hf_energy
        self._label = label
        print(f"Container labeled ({self.label}):\n- parameter: {self.molecule_param}\n"
              f"- exp. value: {self.measured_value}\n- FCI value: {self.fci_value}")

    # accessors (the reference exposes both the get_* methods and the read-only properties below)
    def get_molecule_parameter(self) -> float:
        return self._molecular_parameter

    def set_molecule_parameter(self, m_param: float):
        self._molecular_parameter = m_param

    def get_measured_eigenvalue(self) -> float:
        values = getattr(self, "_optimized_exp_values", None)
        return self._measured_eigenvalue if values is None else statistics.mean(values)

    def get_measured_standard_deviation(self) -> float:
        values = getattr(self, "_optimized_exp_values", None)
        if values is None:
            return getattr(self, "_measured_standard_deviation", 0)
        return statistics.stdev(values) if len(values) > 1 else 0

    def get_fci_value(self) -> float:
        return _or_nan(self._fci_value)

    def get_hf_value(self) -> float:
        return _or_nan(getattr(self, "_hf_value", None))

    def get_label(self) -> str:
        return self._label

    molecule_param, measured_value, measured_std, fci_value, hf_value, label = (
        property(getter) for getter in (get_molecule_parameter, get_measured_eigenvalue, get_measured_standard_deviation,
                                        get_fci_value, get_hf_value, get_label))


class IMeasurementCollection:
    """The sweep of ``calculate_minimisation``: every noise model x every molecule parameter of one ansatz (reference
    :54-95).  ONE ansatz object is re-parameterised and re-yielded per point, as in the reference: consumers must use (or
    deep-copy) it before advancing the generator."""

    def __init__(self, w: IWaveFunction, p_space: List[float], n_space: List[INoiseModel]):
        self.container = []
        self._wave_function_ID = w
        self.parameter_space = p_space
        self.noise_model_space = n_space

    def _bind(self, value: float) -> IWaveFunction:
        wave_function = self._wave_function_ID
        layout = wave_function.molecule_parameters
        for key in layout:
            layout.dict[key] = value
        wave_function.update_molecule(layout)
        return wave_function

    def _get_param_iter(self, wave_func: IWaveFunction) -> Iterator[IWaveFunction]:
        assert wave_func is self._wave_function_ID
        return (self._bind(value) for value in self.parameter_space)

    def get_wave_functions(self) -> Iterator[Tuple[IWaveFunction, str]]:
        for channel in (self.noise_model_space or [None]):
            for value in self.parameter_space:
                bound = self._bind(value)
                if channel is None:
                    yield bound, "No Noise Model"
                else:
                    yield INoiseWrapper(bound, channel), channel.get_description()

    def __iter__(self):
        return iter(self.container)


class IHistogramCollection:
    """Repeated estimates of one quantity (reference :98-114; note ``get_data`` / ``get_mean`` / ``get_std`` are
    properties there)."""

    def __init__(self, measurements: List[float]):
        self._measurements = measurements
        self._mean = float(np.mean(measurements))

    get_data = property(lambda self: self._measurements)
    get_mean = property(lambda self: self._mean)
    get_std = property(lambda self: float(np.std(self._measurements)))


class ISimilarityCollection:
    """Noisy vs mitigated histograms of one experiment (reference :117-127)."""

    def __init__(self, data_noisy: List[float], data_mitigated: List[float], mu_ideal: float, settings: str):
        self.data_noisy, self.data_mitigated = IHistogramCollection(data_noisy), IHistogramCollection(data_mitigated)
        self.mu_ideal = mu_ideal
        self._settings = settings

    @staticmethod
    def get_settings(probability: float, measure_count: int, identifier_count: int) -> str:
        return "P{}_R{}_C{}".format(str(probability), measure_count, identifier_count)
