"""XML (de)serialisation of a RigidBodyIntegrator, node-for-node what the reference's SerializationProxy writes
(serialization/src/RigidBodyIntegratorProxy.cpp:43-67): properties version=1, stepSize, constraintTolerance and a
child node `bodyIndices` holding one `bodyIndex` child per atom with an integer property `index`.  As in the
reference, rotationMode and computeRefinedEnergies are NOT persisted.  The element layout follows OpenMM's
XmlSerializer (properties are XML attributes; the root element is named after the object, with a `type` attribute)."""
from __future__ import annotations

import xml.etree.ElementTree as ET

from ._lib import OpenMMException
from .integrator import RigidBodyIntegrator


def serialize(integrator, root_name="RigidBodyIntegrator"):
    root = ET.Element(root_name, {"type": "RigidBodyIntegrator", "version": "1",
                                  "stepSize": repr(float(integrator.getStepSize())),
                                  "constraintTolerance": repr(float(integrator.getConstraintTolerance()))})
    indices = ET.SubElement(root, "bodyIndices")
    for index in integrator.getBodyIndices():
        ET.SubElement(indices, "bodyIndex", {"index": str(int(index))})
    return ET.tostring(root, encoding="unicode")


def deserialize(xml):
    root = ET.fromstring(xml)
    if int(root.attrib.get("version", "0")) != 1:
        raise OpenMMException("Unsupported version number")
    node = root.find("bodyIndices")
    if node is None:
        raise OpenMMException("Missing bodyIndices node")
    indices = [int(child.attrib["index"]) for child in node.findall("bodyIndex")]
    integrator = RigidBodyIntegrator(float(root.attrib["stepSize"]), indices)
    integrator.setConstraintTolerance(float(root.attrib["constraintTolerance"]))
    return integrator
