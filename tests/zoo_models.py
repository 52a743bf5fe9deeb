"""Feature-coverage models ("zoo"): every joint type, geom/contact primitive, actuator shortcut, sensor and option branch
the MJCF subset accepts, so that parity tests exercise code paths the five BASELINE models do not."""

ZOO_A = """
<mujoco model="zoo_a">
  <compiler angle="radian"/>
  <option timestep="0.004" impratio="2"/>
  <default>
    <geom friction="0.8 0.01 0.001" solref="0.015 1.1" solimp="0.85 0.96 0.002 0.4 3"/>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" margin="0.01" gap="0.002"/>
    <body name="boxy" pos="0 0 0.32" euler="0.1 0.2 0.3">
      <freejoint name="boxroot"/>
      <geom name="box" type="box" size="0.15 0.1 0.2" density="400"/>
      <site name="top" pos="0 0 0.2" euler="0 0.3 0"/>
      <site name="box_sole" type="box" pos="0 0 -0.2" size="0.16 0.11 0.03"/>
      <body name="arm" pos="0.15 0 0.1">
        <joint name="shoulder" type="ball" pos="0 0 0" stiffness="3" damping="0.4" range="0 0.6" margin="0.02"/>
        <geom name="arm" type="capsule" fromto="0 0 0 0.3 0 0" size="0.04"/>
        <body name="hand" pos="0.3 0 0">
          <joint name="wrist" type="hinge" axis="0 1 0" range="-1 1" damping="0.05" armature="0.01" stiffness="1" springref="0.2"/>
          <joint name="extend" type="slide" axis="1 0 0" range="-0.05 0.1" damping="1" margin="0.01"/>
          <geom name="hand" type="sphere" pos="0.08 0 0" size="0.06" condim="1"/>
          <site name="tip" pos="0.14 0 0"/>
        </body>
      </body>
    </body>
    <body name="ball" pos="0.6 0.1 0.12">
      <freejoint name="ballroot"/>
      <geom name="ball" type="sphere" size="0.1" priority="1" friction="0.3 0.02 0.002"/>
      <site name="ball_skin" type="sphere" size="0.13"/>
    </body>
    <body name="rod" pos="-0.5 0 0.06" euler="0 1.5707963 0">
      <freejoint name="rodroot"/>
      <geom name="rod" type="capsule" size="0.05 0.2" solmix="3"/>
      <site name="rod_skin" type="capsule" size="0.07 0.2"/>
      <site name="rod_tip" type="sphere" pos="0 0 0.2" size="0.08"/>
      <inertial pos="0 0 0.01" mass="1.5" fullinertia="0.03 0.03 0.004 0.0005 0 0"/>
    </body>
  </worldbody>
  <contact><exclude body1="boxy" body2="hand"/><exclude body1="boxy" body2="ball"/><exclude body1="boxy" body2="rod"/></contact>
  <actuator>
    <position name="wrist_pos" joint="wrist" kp="4" kv="0.3" ctrlrange="-1 1"/>
    <velocity name="extend_vel" joint="extend" kv="2" forcerange="-3 3"/>
    <general name="wrist_gen" joint="wrist" gear="0.5" gaintype="affine" gainprm="1 0.2 0.1" biastype="affine" biasprm="0.05 -0.3 -0.02"/>
  </actuator>
  <sensor>
    <jointpos joint="wrist"/> <jointvel joint="extend"/>
    <actuatorpos actuator="wrist_pos"/> <actuatorvel actuator="extend_vel"/> <actuatorfrc actuator="wrist_gen"/>
    <framepos objtype="site" objname="tip"/> <framequat objtype="site" objname="top"/> <framepos objtype="body" objname="hand"/>
    <framepos objtype="xbody" objname="arm"/> <framepos objtype="geom" objname="arm"/>
    <framelinvel objtype="site" objname="tip"/> <frameangvel objtype="body" objname="hand"/>
    <velocimeter site="tip"/> <gyro site="top"/> <accelerometer site="tip"/> <accelerometer site="top"/>
    <subtreecom body="boxy"/> <subtreelinvel body="boxy"/> <subtreelinvel body="arm"/> <clock/>
    <touch site="ball_skin"/> <touch site="rod_skin"/> <touch site="rod_tip"/> <touch site="box_sole"/> <touch site="tip"/>
    <force site="tip"/> <torque site="tip"/> <force site="top"/> <torque site="top"/> <force site="ball_skin"/> <torque site="rod_tip"/>
    <framexaxis objtype="site" objname="tip"/> <frameyaxis objtype="body" objname="hand"/> <framezaxis objtype="geom" objname="arm"/>
    <ballquat joint="shoulder"/> <ballangvel joint="shoulder"/> <jointactuatorfrc joint="wrist"/> <jointactuatorfrc joint="extend"/>
    <framelinacc objtype="site" objname="tip"/> <frameangacc objtype="body" objname="hand"/> <framelinacc objtype="geom" objname="ball"/>
  </sensor>
</mujoco>
"""

# CG solver, RK4 with contacts, capsule-capsule and sphere-capsule pairs, frictionless pair, xfrc/qfrc applied in tests
ZOO_B = """
<mujoco model="zoo_b">
  <compiler angle="degree"/>
  <option timestep="0.003" integrator="RK4" solver="CG" tolerance="1e-10" iterations="200">
    <flag warmstart="disable"/>
  </option>
  <worldbody>
    <geom type="plane" size="3 3 0.1" condim="1"/>
    <body name="a" pos="0 0 0.25" euler="0 80 10">
      <freejoint/>
      <geom name="ca" type="capsule" size="0.06 0.25"/>
      <body name="a2" pos="0 0 0.3">
        <joint name="elbow" type="hinge" axis="1 0 0" range="-60 60" damping="0.02"/>
        <geom name="ca2" type="capsule" fromto="0 0 0 0 0 0.3" size="0.05"/>
      </body>
    </body>
    <body name="b" pos="0.05 0.1 0.42" euler="70 0 0">
      <freejoint/>
      <geom name="cb" type="capsule" size="0.05 0.2" condim="1"/>
    </body>
    <body name="s" pos="-0.1 -0.05 0.6">
      <freejoint/>
      <geom name="sph" type="sphere" size="0.08"/>
    </body>
  </worldbody>
  <actuator><motor joint="elbow" gear="2" ctrlrange="-1 1"/></actuator>
  <sensor><subtreelinvel body="a"/><framelinvel objtype="geom" objname="sph"/></sensor>
</mujoco>
"""

# a user-supplied model that is NOT compiled into the library: exercises the run-time specialisation (csrc/ox_jit.cpp)
HOPPER = """
<mujoco model="hopper_user">
  <compiler angle="radian"/>
  <option timestep="0.004"/>
  <default>
    <joint armature="0.02" damping="1" limited="true"/>
    <geom friction="0.9 0.1 0.1" solref="0.02 1" solimp="0.9 0.95 0.001"/>
    <motor ctrllimited="true" ctrlrange="-1 1"/>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="20 20 0.1"/>
    <body name="torso" pos="0 0 1.25">
      <joint name="rootx" type="slide" axis="1 0 0" armature="0" damping="0" limited="false"/>
      <joint name="rootz" type="slide" axis="0 0 1" armature="0" damping="0" limited="false"/>
      <joint name="rooty" type="hinge" axis="0 1 0" armature="0" damping="0" limited="false"/>
      <geom name="torso" type="capsule" fromto="0 0 -0.2 0 0 0.2" size="0.05"/>
      <body name="thigh" pos="0 0 -0.2">
        <joint name="thigh" type="hinge" axis="0 -1 0" range="-2.6 0"/>
        <geom name="thigh" type="capsule" fromto="0 0 0 0 0 -0.45" size="0.05"/>
        <body name="leg" pos="0 0 -0.45">
          <joint name="leg" type="hinge" axis="0 -1 0" range="-2.6 0"/>
          <geom name="leg" type="capsule" fromto="0 0 0 0 0 -0.5" size="0.04"/>
          <body name="foot" pos="0 0 -0.5">
            <joint name="foot" type="hinge" axis="0 -1 0" range="-0.78 0.78"/>
            <geom name="foot" type="capsule" fromto="-0.13 0 -0.06 0.26 0 -0.06" size="0.06" friction="2 0.1 0.1"/>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
    <motor joint="thigh" gear="200"/> <motor joint="leg" gear="200"/> <motor joint="foot" gear="100"/>
  </actuator>
  <sensor><subtreecom body="torso"/><jointvel joint="foot"/></sensor>
</mujoco>
"""

# N4 / N3: implicitfast integrator, stateful actuators (integrator with actrange, filter, filterexact), velocity-dependent
# actuator terms (what implicitfast differentiates), a force-clamped velocity servo, contacts and joint limits
ZOO_C = """
<mujoco model="zoo_c">
  <compiler angle="radian"/>
  <option timestep="0.005" integrator="implicitfast"/>
  <default><joint damping="0.3" armature="0.01"/><geom friction="0.7 0.01 0.001"/></default>
  <worldbody>
    <geom type="plane" size="3 3 0.1"/>
    <body name="base" pos="0 0 0.45">
      <joint name="slide_x" type="slide" axis="1 0 0" damping="2"/>
      <joint name="slide_z" type="slide" axis="0 0 1" range="-0.4 0.6" limited="true"/>
      <geom name="base" type="box" size="0.12 0.08 0.05" density="800"/>
      <body name="upper" pos="0 0 -0.05">
        <joint name="hip" type="hinge" axis="0 1 0" range="-1.2 1.2" limited="true"/>
        <geom name="upper" type="capsule" fromto="0 0 0 0 0 -0.25" size="0.03"/>
        <body name="lower" pos="0 0 -0.25">
          <joint name="knee" type="hinge" axis="0 1 0" range="-2 0.1" limited="true" stiffness="2" springref="-0.4"/>
          <geom name="lower" type="capsule" fromto="0 0 0 0 0 -0.22" size="0.025"/>
          <geom name="toe" type="sphere" pos="0 0 -0.22" size="0.035"/>
        </body>
      </body>
    </body>
  </worldbody>
  <contact><exclude body1="base" body2="lower"/></contact>
  <actuator>
    <general name="hip_int" joint="hip" gear="3" dyntype="integrator" actlimited="true" actrange="-0.6 0.6" ctrlrange="-1 1" ctrllimited="true"
             gaintype="fixed" gainprm="4" biastype="affine" biasprm="0 -4 -0.4"/>
    <general name="knee_filt" joint="knee" gear="2" dyntype="filter" dynprm="0.03" ctrlrange="-1 1" ctrllimited="true"/>
    <general name="x_fexact" joint="slide_x" gear="5" dyntype="filterexact" dynprm="0.02" gaintype="affine" gainprm="1 0 -0.2"/>
    <velocity name="knee_vel" joint="knee" kv="1.5" forcerange="-0.8 0.8"/>
    <position name="hip_pos" joint="hip" kp="3" kv="0.5"/>
  </actuator>
  <sensor><actuatorfrc actuator="hip_int"/><actuatorfrc actuator="x_fexact"/><jointvel joint="knee"/></sensor>
</mujoco>
"""

# RK4 with activation states (the act component of the Runge-Kutta state vector), no contacts
ZOO_D = """
<mujoco model="zoo_d">
  <compiler angle="radian"/>
  <option timestep="0.004" integrator="RK4"/>
  <worldbody>
    <body name="l1" pos="0 0 1">
      <joint name="j1" type="hinge" axis="0 1 0" damping="0.05"/>
      <geom type="capsule" fromto="0 0 0 0 0 -0.4" size="0.03" contype="0" conaffinity="0"/>
      <body name="l2" pos="0 0 -0.4">
        <joint name="j2" type="hinge" axis="0 1 0" range="-2 2" limited="true"/>
        <geom type="capsule" fromto="0 0 0 0 0 -0.3" size="0.025" contype="0" conaffinity="0"/>
      </body>
    </body>
  </worldbody>
  <actuator>
    <general name="a1" joint="j1" dyntype="integrator" actlimited="true" actrange="-0.3 0.3" gainprm="2"/>
    <general name="a2" joint="j2" dyntype="filterexact" dynprm="0.05" gainprm="1.5" biastype="affine" biasprm="0 0 -0.1"/>
  </actuator>
</mujoco>
"""

# N3: equality constraints (connect between two moving bodies and to a mocap body; joint coupling with a quadratic polynomial, one
# of them switched off in the model), a mocap body that carries a colliding geom, contacts on top (src/physics.rs:147-170)
ZOO_E = """
<mujoco model="zoo_e">
  <compiler angle="radian"/>
  <option timestep="0.004"/>
  <default><joint damping="0.1" armature="0.005"/><geom friction="0.8 0.01 0.001"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <body name="hand" mocap="true" pos="0.72 0.02 0.95" quat="0.98 0 0.2 0">
      <geom name="paddle" type="capsule" fromto="-0.2 0 -0.25 0.2 0 -0.25" size="0.04"/>
    </body>
    <body name="upper" pos="0 0 1">
      <joint name="sh" type="hinge" axis="0 1 0"/>
      <geom name="upper" type="capsule" fromto="0 0 0 0.4 0 0" size="0.03"/>
      <body name="fore" pos="0.4 0 0">
        <joint name="el" type="hinge" axis="0 1 0" range="-2 2" limited="true"/>
        <geom name="fore" type="capsule" fromto="0 0 0 0.3 0 0" size="0.025"/>
        <site name="wrist" pos="0.05 0 0" euler="0.2 0 0.1"/>
      </body>
      <site name="shoulder_ft" pos="0 0 0"/>
    </body>
    <body name="gear1" pos="-0.5 0 0.5">
      <joint name="g1" type="hinge" axis="1 0 0"/>
      <geom type="capsule" fromto="0 0 0 0 0.2 0" size="0.03" contype="0" conaffinity="0"/>
      <body name="gear2" pos="0 0.2 0">
        <joint name="g2" type="slide" axis="0 1 0" stiffness="5"/>
        <geom type="sphere" size="0.05" contype="0" conaffinity="0"/>
      </body>
    </body>
    <body name="ball" pos="0.7 0 0.85">
      <freejoint name="ballroot"/>
      <geom name="ball" type="sphere" size="0.07" density="600"/>
    </body>
    <body name="puck" pos="0.2 0.4 0.06">
      <freejoint name="puckroot"/>
      <geom name="puck" type="sphere" size="0.06"/>
    </body>
  </worldbody>
  <contact><exclude body1="fore" body2="ball"/></contact>
  <equality>
    <connect name="hook" body1="fore" body2="hand" anchor="0.3 0 0" solref="0.01 1"/>
    <connect name="tether" body1="ball" body2="fore" anchor="0 0 0.15" solimp="0.8 0.9 0.01 0.5 2"/>
    <joint name="couple" joint1="g1" joint2="g2" polycoef="0.05 1.5 -2 0 0"/>
    <joint name="hold" joint1="sh" polycoef="0.2 0 0 0 0" active="false"/>
    <connect name="pin" body1="puck" anchor="0 0 0.3" active="false"/>
  </equality>
  <actuator><motor joint="sh" gear="3"/><motor joint="g2" gear="2"/></actuator>
  <sensor><jointpos joint="g1"/><framepos objtype="body" objname="ball"/><framepos objtype="body" objname="hand"/>
    <force site="wrist"/><torque site="wrist"/><force site="shoulder_ft"/><torque site="shoulder_ft"/></sensor>
</mujoco>
"""

# N4: the dual solvers. zoo_f = PGS (projected Gauss-Seidel on the constraint forces) followed by the noslip pass, with contacts
# (pyramidal and frictionless), a joint limit and an equality row in the same problem; zoo_g = Newton followed by noslip.
ZOO_F = """
<mujoco model="zoo_f">
  <compiler angle="radian"/>
  <option timestep="0.004" solver="PGS" iterations="60" tolerance="1e-10" noslip_iterations="4" noslip_tolerance="1e-8"/>
  <default><geom friction="1.2 0.01 0.001"/><joint damping="0.05"/></default>
  <worldbody>
    <geom name="ramp" type="plane" size="3 3 0.1" euler="0 0.25 0"/>
    <body name="crate" pos="0 0 0.25" euler="0 0.25 0.1">
      <freejoint name="crateroot"/>
      <geom name="crate" type="box" size="0.12 0.1 0.08" density="500"/>
      <body name="flag" pos="0 0 0.08">
        <joint name="mast" type="hinge" axis="0 1 0" range="-0.5 0.5" limited="true"/>
        <geom name="mast" type="capsule" fromto="0 0 0 0 0 0.25" size="0.015"/>
      </body>
    </body>
    <body name="roller" pos="0.5 0.2 0.2">
      <freejoint name="rollerroot"/>
      <geom name="roller" type="capsule" size="0.06 0.15" euler="1.5707963 0 0" friction="0.4 0.01 0.001"/>
    </body>
    <body name="marble" pos="-0.4 -0.2 0.1">
      <freejoint name="marbleroot"/>
      <geom name="marble" type="sphere" size="0.05" condim="1"/>
    </body>
  </worldbody>
  <contact><exclude body1="crate" body2="roller"/><exclude body1="crate" body2="marble"/></contact>
  <equality><connect name="leash" body1="marble" body2="crate" anchor="0 0 0.25" solref="0.03 1"/></equality>
  <actuator><motor joint="mast" gear="0.5"/></actuator>
  <sensor><framepos objtype="body" objname="crate"/><jointpos joint="mast"/></sensor>
</mujoco>
"""

ZOO_G = HOPPER.replace('model="hopper_user"', 'model="zoo_g"').replace('<option timestep="0.004"/>', '<option timestep="0.004" noslip_iterations="3"/>')

# N3: fixed tendons - a limited tendon coupling two hinges, a spring tendon with a dead band and damping over a hinge and a slide,
# joint limits and contacts in the same problem, tendonpos / tendonvel sensors
ZOO_H = """
<mujoco model="zoo_h">
  <compiler angle="radian"/>
  <option timestep="0.004"/>
  <default><joint damping="0.05" armature="0.005"/><tendon solreflimit="0.015 1"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <body name="base" pos="0 0 0.1">
      <joint name="lift" type="slide" axis="0 0 1" range="-0.03 0.3" limited="true"/>
      <geom name="base" type="sphere" size="0.06"/>
      <body name="f1" pos="0 0 0">
        <joint name="k1" type="hinge" axis="0 1 0" range="-1.5 1.5" limited="true"/>
        <geom name="f1" type="capsule" fromto="0 0 0 0.25 0 0" size="0.025"/>
        <body name="f2" pos="0.25 0 0">
          <joint name="k2" type="hinge" axis="0 1 0"/>
          <geom name="f2" type="capsule" fromto="0 0 0 0.2 0 0" size="0.02"/>
          <body name="f3" pos="0.2 0 0">
            <joint name="k3" type="hinge" axis="0 1 0"/>
            <geom name="f3" type="capsule" fromto="0 0 0 0.15 0 0" size="0.018"/>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <tendon>
    <fixed name="curl" limited="true" range="-0.6 0.9" margin="0.02"><joint joint="k2" coef="1"/><joint joint="k3" coef="-0.7"/></fixed>
    <fixed name="spring" stiffness="12" damping="0.4" springlength="-0.1 0.15"><joint joint="k1" coef="0.5"/><joint joint="lift" coef="2"/></fixed>
    <fixed name="soft" stiffness="3" range="-2 2"><joint joint="k1" coef="1"/><joint joint="k2" coef="1"/><joint joint="k3" coef="1"/></fixed>
  </tendon>
  <actuator><motor joint="k1" gear="2"/><motor joint="k3" gear="0.6"/>
    <motor name="pull" tendon="curl" gear="0.4"/><position name="servo" tendon="soft" kp="3" kv="0.2" gear="1.5" forcerange="-2 2"/></actuator>
  <sensor><tendonpos tendon="curl"/><tendonvel tendon="spring"/><tendonpos tendon="soft"/><jointpos joint="k2"/>
    <actuatorpos actuator="servo"/><actuatorvel actuator="servo"/><actuatorfrc actuator="pull"/></sensor>
</mujoco>
"""

# N3: box narrowphase beyond box-plane - spheres and capsules against free and static boxes (faces, edges, corners, corners), on top of the plane pairs; the box-box pair (slab / step) is masked out - that narrowphase is outside the supported set. zoo_f's crate / roller / marble exclusions are not needed any more.
ZOO_I = """
<mujoco model="zoo_i">
  <compiler angle="radian"/>
  <option timestep="0.003" tolerance="1e-13"/>
  <default><geom friction="0.9 0.01 0.001"/><default class="mover"><geom contype="5"/></default></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <geom name="step" type="box" pos="0.5 0 0.1" size="0.25 0.3 0.1" euler="0 0 0.3" contype="2" conaffinity="4"/>
    <body name="slab" pos="-0.4 0 0.08" euler="0.05 0.1 0.2">
      <freejoint/>
      <geom name="slab" type="box" size="0.3 0.25 0.06" density="300"/>
    </body>
    <body name="ball1" pos="-0.45 0.05 0.3"><freejoint/><geom name="ball1" class="mover" type="sphere" size="0.07"/></body>
    <body name="ball2" pos="0.45 0.28 0.35"><freejoint/><geom name="ball2" class="mover" type="sphere" size="0.06" condim="1"/></body>
    <body name="rod1" pos="-0.3 -0.1 0.35" euler="0.4 1.2 0"><freejoint/><geom name="rod1" class="mover" type="capsule" size="0.04 0.18"/></body>
    <body name="rod2" pos="0.62 -0.05 0.32" euler="1.5707963 0.1 0.3"><freejoint/><geom name="rod2" class="mover" type="capsule" size="0.035 0.22"/></body>
    <body name="rod3" pos="0.3 -0.32 0.42" euler="0.2 0.1 0"><freejoint/><geom name="rod3" class="mover" type="capsule" size="0.03 0.15"/></body>
  </worldbody>
  <sensor><framepos objtype="body" objname="ball1"/><framepos objtype="body" objname="rod2"/></sensor>
</mujoco>
"""

# N3: condim 4 (torsional friction) and condim 6 (torsional + rolling): spinning / rolling bodies on the floor and on each other,
# with touch / force / torque sensors reading the contact wrench (the torsional and rolling rows are pure torques)
ZOO_J = """
<mujoco model="zoo_j">
  <compiler angle="radian"/>
  <option timestep="0.003" tolerance="1e-13" impratio="3"/>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1" condim="3" friction="1 0.05 0.01"/>
    <body name="top" pos="0 0 0.11">
      <freejoint name="toproot"/>
      <geom name="top" type="sphere" size="0.1" condim="4" friction="0.8 0.03 0.0001"/>
      <site name="topsite" pos="0 0 -0.05"/>
    </body>
    <body name="wheel" pos="0.5 0 0.09">
      <freejoint name="wheelroot"/>
      <geom name="wheel" type="sphere" size="0.08" condim="6" friction="0.7 0.02 0.004"/>
      <site name="wheelskin" type="sphere" size="0.1"/>
    </body>
    <body name="log" pos="-0.5 0 0.07" euler="1.5707963 0 0.4">
      <freejoint name="logroot"/>
      <geom name="log" type="capsule" size="0.06 0.2" condim="6" friction="0.6 0.01 0.002" priority="1"/>
      <site name="logsite" pos="0 0 0.1"/>
    </body>
    <body name="rider" pos="-0.5 0 0.21">
      <freejoint name="riderroot"/>
      <geom name="rider" type="sphere" size="0.07" condim="4" friction="0.9 0.04 0.0001"/>
    </body>
  </worldbody>
  <sensor>
    <touch site="wheelskin"/><force site="topsite"/><torque site="topsite"/><torque site="logsite"/>
    <frameangvel objtype="body" objname="top"/><frameangvel objtype="body" objname="wheel"/>
  </sensor>
</mujoco>
"""

# N3: dry joint friction (frictionloss): Huber-cost rows on hinges, a slide and a ball joint, Newton here and PGS + noslip in
# zoo_l, together with limits and a contact
ZOO_K = """
<mujoco model="zoo_k">
  <compiler angle="radian"/>
  <option timestep="0.004" tolerance="1e-13"/>
  <default><joint damping="0.02" armature="0.002"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <body name="cart" pos="0 0 0.5">
      <joint name="rail" type="slide" axis="1 0 0" frictionloss="1.5" range="-1 1" limited="true"/>
      <geom name="cart" type="box" size="0.1 0.08 0.05" density="600" contype="0" conaffinity="0"/>
      <body name="arm1" pos="0 0 -0.05">
        <joint name="a1" type="hinge" axis="0 1 0" frictionloss="0.15" solreffriction="0.01 1" solimpfriction="0.8 0.9 0.001 0.5 2"/>
        <geom name="arm1" type="capsule" fromto="0 0 0 0 0 -0.25" size="0.02"/>
        <body name="arm2" pos="0 0 -0.25">
          <joint name="a2" type="ball" frictionloss="0.05"/>
          <geom name="arm2" type="capsule" fromto="0 0 0 0 0 -0.2" size="0.018"/>
          <geom name="bob" type="sphere" pos="0 0 -0.2" size="0.04"/>
        </body>
      </body>
    </body>
    <body name="free" pos="0.4 0.1 0.3"><freejoint/><geom name="free" type="sphere" size="0.06"/>
      <body name="flap" pos="0 0 0.06"><joint name="fl" type="hinge" axis="1 0 0" frictionloss="0.02" range="-1 1" limited="true"/>
        <geom name="flap" type="capsule" fromto="0 0 0 0 0.12 0" size="0.012"/></body></body>
  </worldbody>
  <actuator><motor joint="rail" gear="4"/><motor joint="a1" gear="0.5"/><motor joint="fl" gear="0.05"/></actuator>
  <sensor><jointpos joint="rail"/><jointvel joint="a1"/><jointpos joint="fl"/></sensor>
</mujoco>
"""

ZOO_L = ZOO_K.replace('model="zoo_k"', 'model="zoo_l"').replace('<option timestep="0.004" tolerance="1e-13"/>',
                                                              '<option timestep="0.004" solver="PGS" iterations="80" tolerance="1e-12" noslip_iterations="3" noslip_tolerance="1e-9"/>')

# N3: elliptic friction cones (Newton here, CG in zoo_n): condim 3, 4 and 6 contacts in the same problem, sliding, sticking and
# spinning bodies, force / torque / touch sensors reading the cone's force directly
ZOO_M = ZOO_J.replace('model="zoo_j"', 'model="zoo_m"').replace('<option timestep="0.003" tolerance="1e-13" impratio="3"/>',
                                                              '<option timestep="0.003" tolerance="1e-13" impratio="3" cone="elliptic"/>')
ZOO_N = HOPPER.replace('model="hopper_user"', 'model="zoo_n"').replace('<option timestep="0.004"/>',
                                                                     '<option timestep="0.004" cone="elliptic" solver="CG" tolerance="1e-12" iterations="300"/>')

# N3: weld equality - a free body welded to a mocap body (the standard mocap drive: position AND orientation), a weld between two
# moving bodies with an explicit relpose and torquescale, one weld inactive in the model; force / torque sensors read the weld wrench
ZOO_O = """
<mujoco model="zoo_o">
  <compiler angle="radian"/>
  <option timestep="0.004" tolerance="1e-13"/>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <body name="hand" mocap="true" pos="0.3 0.1 0.6" quat="0.95 0.1 0.25 0.05"/>
    <body name="tool" pos="0.3 0.1 0.6" quat="0.95 0.1 0.25 0.05">
      <freejoint name="toolroot"/>
      <geom name="tool" type="capsule" fromto="0 0 0 0.25 0 0" size="0.03"/>
      <site name="grip" pos="0.05 0 0"/>
      <body name="tip" pos="0.25 0 0">
        <joint name="wrist" type="hinge" axis="0 0 1" damping="0.05"/>
        <geom name="tip" type="capsule" fromto="0 0 0 0.12 0 0" size="0.02"/>
      </body>
    </body>
    <body name="base" pos="-0.4 0 0.4">
      <joint name="b1" type="hinge" axis="0 1 0" damping="0.1"/>
      <geom name="base" type="capsule" fromto="0 0 0 0 0 -0.3" size="0.03"/>
      <site name="basesite" pos="0 0 -0.1"/>
    </body>
    <body name="rider" pos="-0.4 0.1 0.15">
      <freejoint name="riderroot"/>
      <geom name="rider" type="box" size="0.06 0.04 0.03" density="400"/>
    </body>
    <body name="pebble" pos="0.8 0 0.0495"><freejoint name="pebbleroot"/><geom name="pebble" type="sphere" size="0.05"/></body>
  </worldbody>
  <equality>
    <weld name="grasp" body1="tool" body2="hand" solref="0.01 1"/>
    <weld name="saddle" body1="rider" body2="base" relpose="0 0.1 -0.25 0.98 0 0.2 0" anchor="0 0 -0.25" torquescale="0.5"/>
    <weld name="spare" body1="tip" active="false"/>
  </equality>
  <actuator><motor joint="wrist" gear="0.3"/><motor joint="b1" gear="1.5"/></actuator>
  <sensor><force site="grip"/><torque site="grip"/><torque site="basesite"/><framequat objtype="body" objname="tool"/></sensor>
</mujoco>
"""

# N3 / north_star "box contacts": box-box narrowphase - boxes dropped on a static box and on each other (face-face with clipped
# polygons, edge and corner touches), next to box-plane, sphere-box and capsule-box pairs
ZOO_P = """
<mujoco model="zoo_p">
  <compiler angle="radian"/>
  <option timestep="0.003" tolerance="1e-13"/>
  <default><geom friction="0.8 0.01 0.001"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <geom name="table" type="box" pos="0 0 0.15" size="0.4 0.3 0.15" euler="0 0 0.2"/>
    <body name="crate" pos="0.05 0.02 0.42" euler="0.1 0.05 0.4"><freejoint/><geom name="crate" type="box" size="0.12 0.1 0.08" density="500"/></body>
    <body name="brick" pos="0.1 0.0 0.62" euler="0.3 0.2 1.0"><freejoint/><geom name="brick" type="box" size="0.08 0.04 0.03" density="800"/></body>
    <body name="plank" pos="-0.3 0.25 0.4" euler="0.5 0.1 0"><freejoint/><geom name="plank" type="box" size="0.2 0.03 0.02" density="600"/></body>
    <body name="marble" pos="-0.1 -0.1 0.45"><freejoint/><geom name="marble" type="sphere" size="0.04"/></body>
  </worldbody>
  <sensor><framepos objtype="body" objname="crate"/><framepos objtype="body" objname="brick"/></sensor>
</mujoco>
"""

# N3: spatial tendons (straight segments through sites): a limited tendon from the world to a swinging arm's tip through an
# intermediate site, a spring-damper tendon between two free bodies, an actuator pulling on a spatial tendon
ZOO_Q = """
<mujoco model="zoo_q">
  <compiler angle="radian"/>
  <option timestep="0.004" tolerance="1e-13"/>
  <default><joint damping="0.05" armature="0.003"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <site name="anchor" pos="0.1 0 1.1"/>
    <body name="arm" pos="0 0 1">
      <joint name="sh" type="hinge" axis="0 1 0"/>
      <geom name="arm" type="capsule" fromto="0 0 0 0.4 0 0" size="0.03"/>
      <site name="mid" pos="0.2 0 0.05"/>
      <body name="fore" pos="0.4 0 0">
        <joint name="el" type="hinge" axis="0 1 0" range="-2.2 2.2" limited="true"/>
        <joint name="tw" type="hinge" axis="1 0 0"/>
        <geom name="fore" type="capsule" fromto="0 0 0 0.3 0 0" size="0.025"/>
        <site name="tip" pos="0.3 0 0.03"/>
      </body>
    </body>
    <body name="puck" pos="0.5 0.2 0.4"><freejoint/><geom name="puck" type="sphere" size="0.05"/><site name="puck_c" pos="0 0 0.05"/></body>
    <body name="bob" pos="0.6 -0.1 0.25"><freejoint/><geom name="bob" type="sphere" size="0.04"/><site name="bob_c" pos="0.01 0 0"/></body>
    <body name="pebble" pos="-0.6 0.3 0.0495"><freejoint name="pebbleroot"/><geom name="pebble" type="sphere" size="0.05"/></body>
  </worldbody>
  <tendon>
    <spatial name="leash" limited="true" range="0 0.62" solreflimit="0.015 1"><site site="anchor"/><site site="mid"/><site site="tip"/></spatial>
    <spatial name="bungee" stiffness="40" damping="0.8" springlength="0.25"><site site="tip"/><site site="puck_c"/></spatial>
    <spatial name="cord" stiffness="25" range="0.05 0.5"><site site="puck_c"/><site site="bob_c"/></spatial>
  </tendon>
  <actuator><motor name="winch" tendon="leash" gear="-3"/><motor joint="tw" gear="0.2"/></actuator>
  <sensor><tendonpos tendon="leash"/><tendonvel tendon="bungee"/><tendonpos tendon="cord"/><actuatorpos actuator="winch"/></sensor>
</mujoco>
"""

# mj_passive fluid forces (inertia-box model): a three-link swimmer on planar root joints (dm_control swimmer's structure) in a
# dense, viscous medium with wind, next to a tumbling free plate that falls to the floor and a resting pebble; body gravcomp: a
# neutrally buoyant float and a half-compensated tail segment (its hinge is tilted so that gravity has a moment about it)
ZOO_R = """
<mujoco model="zoo_r">
  <compiler angle="radian"/>
  <option timestep="0.003" tolerance="1e-13" density="900" viscosity="0.2" wind="0.3 -0.1 0.05"/>
  <default><joint damping="0.02" armature="0.002"/></default>
  <worldbody>
    <geom name="floor" type="plane" size="3 3 0.1"/>
    <body name="head" pos="0 0 0.6">
      <joint name="rootx" type="slide" axis="1 0 0"/>
      <joint name="rooty" type="slide" axis="0 1 0"/>
      <joint name="rootz" type="hinge" axis="0 0 1"/>
      <geom name="head" type="capsule" fromto="0 0 0 -0.2 0 0" size="0.03" density="1000"/>
      <body name="seg1" pos="-0.2 0 0">
        <joint name="j1" type="hinge" axis="0 0 1" range="-1.6 1.6" limited="true"/>
        <geom name="seg1" type="box" pos="-0.1 0 0" size="0.1 0.015 0.04" density="1000"/>
        <body name="seg2" pos="-0.2 0 0" gravcomp="0.5">
          <joint name="j2" type="hinge" axis="0.1 0 1" range="-1.6 1.6" limited="true"/>
          <geom name="seg2" type="capsule" fromto="0 0 0 -0.2 0 0" size="0.025" density="1000"/>
          <site name="tail" pos="-0.2 0 0"/>
        </body>
      </body>
    </body>
    <body name="plate" pos="0.7 0.3 0.35" euler="0.4 0.2 0.1"><freejoint/><geom name="plate" type="box" size="0.15 0.1 0.01" density="1500"/></body>
    <body name="float" pos="-0.5 -0.6 0.5" gravcomp="1"><freejoint/><geom name="float" type="sphere" size="0.06" density="300"/></body>
    <body name="pebble" pos="0.8 -0.7 0.0495"><freejoint name="pebbleroot"/><geom name="pebble" type="sphere" size="0.05" density="9000"/></body>
  </worldbody>
  <actuator><motor joint="j1" gear="0.8"/><motor joint="j2" gear="0.8"/></actuator>
  <sensor><framepos objtype="site" objname="tail"/><framelinvel objtype="body" objname="plate"/><gyro site="tail"/></sensor>
</mujoco>
"""

ZOO = {"zoo_r": ZOO_R, "zoo_a": ZOO_A, "zoo_b": ZOO_B, "zoo_c": ZOO_C, "zoo_e": ZOO_E, "zoo_f": ZOO_F, "zoo_g": ZOO_G, "zoo_h": ZOO_H, "zoo_i": ZOO_I, "zoo_j": ZOO_J, "zoo_k": ZOO_K, "zoo_l": ZOO_L, "zoo_m": ZOO_M, "zoo_n": ZOO_N, "zoo_o": ZOO_O, "zoo_p": ZOO_P, "zoo_q": ZOO_Q}
NOCONTACT = {"zoo_d": ZOO_D}
