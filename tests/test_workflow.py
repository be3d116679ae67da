"""
The reference's calculator workflow expectations (tests/calculators/test_workflow.py:112-192) for the PME /
P3M calculators of this package, on the CPU here and on CUDA on the GPU box: dtype / device preserved, runs
as Python, as a TorchScript module (``torch.jit.script``), after ``jit.save`` / ``jit.load``, gradients finite.
"""
import io

import pytest
import torch

DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


def cscl_system(device, dtype):
    """CsCl crystal, same as the reference fixture (test_workflow.py:65-75)"""
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], dtype=dtype, device=device)
    charges = torch.tensor([1.0, -1.0], dtype=dtype, device=device).reshape((-1, 1))
    cell = torch.eye(3, dtype=dtype, device=device)
    neighbor_indices = torch.tensor([[0, 1]], dtype=torch.int64, device=device)
    neighbor_distances = torch.tensor([0.8660], dtype=dtype, device=device)
    return charges, cell, positions, neighbor_indices, neighbor_distances


def make(cls_name, device, dtype, potential="coulomb"):
    import torchpme_b200 as tp

    if potential == "coulomb":
        pot = tp.CoulombPotential(smearing=0.1)
    elif potential == "ipl":
        pot = tp.InversePowerLawPotential(exponent=3, smearing=0.1, exclusion_radius=0.5)
    cls = getattr(tp, cls_name)
    calc = cls(potential=pot, mesh_spacing=0.1)
    calc.to(device=device, dtype=dtype)
    return calc


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("cls_name", ["PMECalculator", "P3MCalculator"])
class TestWorkflow:
    def check_operation(self, calculator, device, dtype):
        descriptor = calculator.forward(*cscl_system(device, dtype))
        assert type(descriptor) is torch.Tensor
        return descriptor

    def test_dtype_device(self, cls_name, device, dtype):
        potential = make(cls_name, device, dtype)(*cscl_system(device, dtype))
        assert potential.dtype == dtype and potential.device.type == device

    def test_operation_as_python(self, cls_name, device, dtype):
        self.check_operation(make(cls_name, device, dtype), device, dtype)

    @pytest.mark.parametrize("potential", ["coulomb", "ipl"])
    def test_operation_as_torch_script(self, cls_name, device, dtype, potential):
        calculator = make(cls_name, device, dtype, potential)
        eager = self.check_operation(calculator, device, dtype)
        scripted = torch.jit.script(calculator)
        out = self.check_operation(scripted, device, dtype)
        assert torch.allclose(out, eager)

    def test_save_load(self, cls_name, device, dtype):
        calculator = make(cls_name, device, dtype)
        scripted = torch.jit.script(calculator)
        with io.BytesIO() as buffer:
            torch.jit.save(scripted, buffer)
            buffer.seek(0)
            loaded = torch.jit.load(buffer)
        # the loaded shell still reaches the live calculator through the dispatcher operator
        out = loaded(*cscl_system(device, dtype))
        assert torch.allclose(out, calculator(*cscl_system(device, dtype)))

    def test_not_nan(self, cls_name, device, dtype):
        calculator = make(cls_name, device, dtype)
        system = list(cscl_system(device, dtype))
        for k in (0, 1, 2, 4):
            system[k].requires_grad = True
        energy = calculator.forward(*system).sum()
        for k in (0, 4, 1, 2):      # charges, distances, cell, positions
            assert not torch.isnan(torch.autograd.grad(energy, system[k], retain_graph=True)[0]).any()

    def test_gradients_through_the_scripted_module(self, cls_name, device, dtype):
        calculator = make(cls_name, device, dtype)
        scripted = torch.jit.script(calculator)
        grads = []
        for module in (calculator, scripted):
            system = list(cscl_system(device, dtype))
            system[2].requires_grad = True
            (module.forward(*system) * system[0]).sum().backward()
            grads.append(system[2].grad)
        assert torch.allclose(grads[0], grads[1])
