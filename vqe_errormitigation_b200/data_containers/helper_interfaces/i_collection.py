"""Result containers (mirror of ``src/data_containers/helper_interfaces/i_collection.py``)."""
from __future__ import annotations

from statistics import mean, stdev
from typing import Iterator, List, Tuple

import numpy as np

from .i_noise_wrapper import INoiseModel, INoiseWrapper
from .i_parameter import IParameter
from .i_wave_function import IWaveFunction


class IContainer:
    def __init__(self, m_param: IParameter, e_values: List[float], m_data, label: str = "<input label>"):
        self._molecular_parameter = m_param["r0"]
        self._optimized_exp_values = e_values
        self._fci_value = m_data.fci_energy
        self._hf_value = m_data.hf_energy
        self._label = label
        print(f"Container labeled ({self.label}):\n- parameter: {self.molecule_param}\n- exp. value: {self.measured_value}\n- FCI value: {self.fci_value}")

    def get_molecule_parameter(self) -> float:
        return self._molecular_parameter

    def set_molecule_parameter(self, m_param: float):
        self._molecular_parameter = m_param

    def get_measured_eigenvalue(self) -> float:
        if not hasattr(self, "_optimized_exp_values"):      # files stored with the reference's earlier schema
            return self._measured_eigenvalue                 # (src/classic_minimisation/H2_{bitflip,depolar}_*, H2_*semi*)
        return mean(self._optimized_exp_values)

    def get_measured_standard_deviation(self) -> float:
        if not hasattr(self, "_optimized_exp_values"):
            return getattr(self, "_measured_standard_deviation", 0)
        return stdev(self._optimized_exp_values) if len(self._optimized_exp_values) >= 2 else 0

    def get_fci_value(self) -> float:
        return self._fci_value if self._fci_value is not None else float("NaN")

    def get_hf_value(self) -> float:
        value = getattr(self, "_hf_value", None)
        return value if value is not None else float("NaN")

    def get_label(self) -> str:
        return self._label

    molecule_param = property(get_molecule_parameter)
    measured_value = property(get_measured_eigenvalue)
    measured_std = property(get_measured_standard_deviation)
    fci_value = property(get_fci_value)
    hf_value = property(get_hf_value)
    label = property(get_label)


class IMeasurementCollection:
    def __init__(self, w: IWaveFunction, p_space: List[float], n_space: List[INoiseModel]):
        self.container = list()
        self._wave_function_ID = w
        self.parameter_space = p_space
        self.noise_model_space = n_space

    def get_wave_functions(self) -> Iterator[Tuple[IWaveFunction, str]]:
        """(noise model x molecule parameter) sweep; one ansatz object is mutated and re-yielded per
        parameter (reference :56-78)."""
        if len(self.noise_model_space) == 0:
            for wave_function in self._get_param_iter(self._wave_function_ID):
                yield wave_function, "No Noise Model"
        else:
            for channel in self.noise_model_space:
                for wave_function in self._get_param_iter(self._wave_function_ID):
                    yield INoiseWrapper(wave_function, channel), channel.get_description()

    def _get_param_iter(self, wave_func: IWaveFunction) -> Iterator[IWaveFunction]:
        layout = wave_func.molecule_parameters
        for value in self.parameter_space:
            for key in layout:
                layout.dict[key] = value
            wave_func.update_molecule(layout)
            yield wave_func

    def __iter__(self):
        return iter(self.container)


class IHistogramCollection:
    def __init__(self, measurements: List[float]):
        self._measurements = measurements
        self._mean = float(np.mean(self._measurements))

    @property
    def get_data(self) -> List[float]:
        return self._measurements

    @property
    def get_mean(self) -> float:
        return self._mean

    @property
    def get_std(self) -> float:
        return float(np.std(self._measurements))


class ISimilarityCollection:
    def __init__(self, data_noisy: List[float], data_mitigated: List[float], mu_ideal: float, settings: str):
        self.data_noisy = IHistogramCollection(measurements=data_noisy)
        self.data_mitigated = IHistogramCollection(measurements=data_mitigated)
        self.mu_ideal = mu_ideal
        self._settings = settings

    @staticmethod
    def get_settings(probability: float, measure_count: int, identifier_count: int) -> str:
        return f"P{probability.__str__()}_R{measure_count}_C{identifier_count}"
