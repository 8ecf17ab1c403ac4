# install_shims.py  (import before anything from sigmarl; needs env CICD_TESTING=true)
"""Probe-only: minimal stand-ins for vmas / torchdiffeq / termcolor + permissive stubs for
tensordict / torchrl / matplotlib / cvxpy so /root/reference/sigmarl hot-path modules import on CPU."""
import sys, types, torch

# ---------- permissive stubs ----------
class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith('__'): raise AttributeError(k)
        return _mk(k)
    def __setitem__(cls, k, v): pass
    def __getitem__(cls, k): return _mk('item')
    def __or__(cls, o): return cls
    def __ror__(cls, o): return cls
def _mk(name):
    return _Meta(name, (object,), {'__init__': lambda self,*a,**k: None,
                                   '__call__': lambda self,*a,**k: None,
                                   '__getattr__': lambda self,k: _mk(k)})
class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'): raise AttributeError(k)
        v=_mk(k); setattr(self,k,v); return v
def stub(*names):
    for n in names:
        parts=n.split('.')
        for i in range(1,len(parts)+1):
            nm='.'.join(parts[:i])
            if nm not in sys.modules:
                m=_Stub(nm); m.__path__=[]; sys.modules[nm]=m
stub('matplotlib.pyplot','matplotlib.patches','matplotlib.patheffects','matplotlib.colors','matplotlib.ticker','matplotlib.animation','matplotlib.lines',
     'tensordict.tensordict','tensordict.nn.distributions','torchrl.collectors','torchrl.envs.utils','torchrl.envs.common','torchrl.envs.libs.vmas','torchrl._utils',
     'torchrl.data.utils','torchrl.data.replay_buffers.samplers','torchrl.data.replay_buffers.storages','torchrl.modules','torchrl.objectives','cvxpy','pyglet')
sys.modules['matplotlib'].colormaps=_mk('colormaps')
def _override(cls):
    def deco(f): return f
    return deco

# ---------- termcolor ----------
tc=types.ModuleType('termcolor'); tc.colored=lambda s,*a,**k:s; tc.cprint=lambda s,*a,**k:None; sys.modules['termcolor']=tc

# ---------- torchdiffeq (fixed-grid euler only) ----------
td=types.ModuleType('torchdiffeq')
def odeint(func,y0,t,rtol=None,atol=None,method='euler'):
    assert method=='euler'
    sol=[y0]; y=y0
    for t0,t1 in zip(t[:-1],t[1:]):
        dt=t1-t0; y=y+dt*func(t0,y); sol.append(y)
    return torch.stack(sol)
td.odeint=odeint; sys.modules['torchdiffeq']=td

# ---------- vmas (memory reconstruction of 1.4.3 semantics used by sigmarl) ----------
class TorchUtils:
    @staticmethod
    def where_from_index(env_index,new_value,old_value):
        mask=torch.zeros_like(old_value,dtype=torch.bool); mask[env_index]=True
        return torch.where(mask,new_value,old_value)
class TVO:
    def __init__(self): self._batch_dim=None; self._device=None
    @property
    def batch_dim(self): return self._batch_dim
    @batch_dim.setter
    def batch_dim(self,v): self._batch_dim=v
    @property
    def device(self): return self._device
    @device.setter
    def device(self,v): self._device=v
class Box:
    def __init__(self,length=0.3,width=0.1,hollow=False): self.length=length; self.width=width
def _prop(name):
    def g(self): return getattr(self,'_'+name)
    def s(self,v):
        assert self._batch_dim is not None and v.shape[0]==self._batch_dim
        setattr(self,'_'+name,v.to(self._device))
    return property(g,s)
class EntityState(TVO):
    def __init__(self):
        super().__init__(); self._pos=self._vel=self._rot=self._ang_vel=None
    pos=_prop('pos'); vel=_prop('vel'); rot=_prop('rot'); ang_vel=_prop('ang_vel')
    def _reset(self,env_index):
        for a in ['pos','rot','vel','ang_vel']:
            v=getattr(self,a)
            if v is not None:
                setattr(self,a, torch.zeros_like(v) if env_index is None else TorchUtils.where_from_index(env_index,0,v))
    def _spawn(self,dim_c,dim_p):
        z=lambda d: torch.zeros(self.batch_dim,d,device=self.device,dtype=torch.float32)
        self.pos=z(dim_p); self.vel=z(dim_p); self.rot=z(1); self.ang_vel=z(1)
class AgentState(EntityState):
    def __init__(self): super().__init__(); self._c=self._force=self._torque=None
    c=_prop('c'); force=_prop('force'); torque=_prop('torque')
    def _reset(self,env_index):
        for a in ['c','force','torque']:
            v=getattr(self,a)
            if v is not None:
                setattr(self,a, torch.zeros_like(v) if env_index is None else TorchUtils.where_from_index(env_index,0,v))
        super()._reset(env_index)
    def _spawn(self,dim_c,dim_p):
        self.force=torch.zeros(self.batch_dim,dim_p,device=self.device); self.torque=torch.zeros(self.batch_dim,1,device=self.device)
        super()._spawn(dim_c,dim_p)
class Action(TVO):
    def __init__(self,u_range,u_multiplier,action_size):
        super().__init__(); self.u=None; self.c=None; self.u_range=u_range; self.u_multiplier=u_multiplier; self.action_size=action_size
    def _reset(self,env_index):
        if self.u is not None:
            self.u = torch.zeros_like(self.u) if env_index is None else TorchUtils.where_from_index(env_index,0,self.u)
class Dynamics:
    def __init__(self): self._agent=None
    @property
    def agent(self): return self._agent
    @agent.setter
    def agent(self,v): self._agent=v
    def reset(self,index=None): pass
    def zero_grad(self): pass
    def check_and_process_action(self):
        assert self.agent.action.u.shape[1]>=self.needed_action_size; self.process_action()
class Agent(TVO):
    def __init__(self,name,shape=None,color=None,collide=True,render_action=False,u_range=1.0,u_multiplier=1.0,max_speed=None,dynamics=None,**kw):
        super().__init__(); self.name=name; self.shape=shape; self._color=color; self.max_speed=max_speed; self.u_range=u_range
        self.dynamics=dynamics; dynamics.agent=self; self.action_script=None
        self._action=Action(u_range,u_multiplier,dynamics.needed_action_size); self._state=AgentState()
    @property
    def state(self): return self._state
    @property
    def action(self): return self._action
    def _spawn(self,dim_c,dim_p):
        for o in (self._state,self._action): o.batch_dim=self.batch_dim; o.device=self.device
        self._state._spawn(dim_c,dim_p)
    def _reset(self,env_index):
        self._action._reset(env_index); self.dynamics.reset(env_index); self._state._reset(env_index)
    def _set_state_property(self,prop,entity,new,batch_index):
        if batch_index is None:
            if len(new.shape)>1 and new.shape[0]==self.batch_dim: prop.fset(entity,new)
            else: prop.fset(entity,new.repeat(self.batch_dim,1))
        else:
            value=prop.fget(entity); value[batch_index]=new
    def set_pos(self,v,batch_index): self._set_state_property(EntityState.pos,self.state,v,batch_index)
    def set_vel(self,v,batch_index): self._set_state_property(EntityState.vel,self.state,v,batch_index)
    def set_rot(self,v,batch_index): self._set_state_property(EntityState.rot,self.state,v,batch_index)
class World(TVO):
    def __init__(self,batch_dim,device,dt=0.1,x_semidim=None,y_semidim=None,**kw):
        super().__init__(); self._batch_dim=batch_dim; self._device=device; self._agents=[]; self._dt=dt; self._x_semidim=x_semidim; self._y_semidim=y_semidim
    dt=property(lambda s:s._dt); x_semidim=property(lambda s:s._x_semidim); y_semidim=property(lambda s:s._y_semidim)
    agents=property(lambda s:s._agents); entities=property(lambda s:s._agents); policy_agents=property(lambda s:s._agents)
    def add_agent(self,agent):
        agent.batch_dim=self._batch_dim; agent.device=self._device; agent._spawn(0,2); self._agents.append(agent)
    def reset(self,env_index):
        for e in self.entities: e._reset(env_index)
class BaseScenario:
    def __init__(self): self._world=None
    @property
    def world(self): return self._world
    def env_make_world(self,batch_dim,device,**kw): self._world=self.make_world(batch_dim,device,**kw); return self._world
    def env_reset_world_at(self,env_index): self.world.reset(env_index); self.reset_world_at(env_index)
    def env_process_action(self,agent): self.process_action(agent); agent.dynamics.check_and_process_action()
    def process_action(self,agent): pass
    def pre_step(self): pass
    def post_step(self): pass
    def info(self,agent): return {}
class Environment:
    """vmas.simulator.environment.Environment (subset): step/reset/reset_at/get_from_scenario ordering."""
    def __init__(self,scenario,num_envs,device='cpu',max_steps=None,seed=None,**kw):
        self.scenario=scenario; self.num_envs=num_envs; self.device=device; self.max_steps=max_steps
        self.world=scenario.env_make_world(num_envs,device,**kw); self.agents=self.world.policy_agents; self.n_agents=len(self.agents)
        if seed is not None: torch.manual_seed(seed)
        self.reset()
    def reset(self,return_observations=True):
        self.scenario.env_reset_world_at(None); self.steps=torch.zeros(self.num_envs)
        return self.get_from_scenario(return_observations,False,False,False)
    def reset_at(self,index,return_observations=True):
        self.scenario.env_reset_world_at(index); self.steps[index]=0
        return self.get_from_scenario(return_observations,False,False,False)
    def get_from_scenario(self,get_observations,get_rewards,get_infos,get_dones):
        obs,rews,infos=[],[],[]
        for agent in self.agents:
            if get_rewards: rews.append(self.scenario.reward(agent).clone())
            if get_observations: obs.append(self.scenario.observation(agent).clone())
            if get_infos: infos.append({k:(v.clone() if torch.is_tensor(v) else v) for k,v in self.scenario.info(agent).items()})
        dones=None
        if get_dones:
            dones=self.scenario.done().clone()
            if self.max_steps is not None: dones=dones | (self.steps>=self.max_steps)
        return obs,rews,dones,infos
    def step(self,actions):
        for a,agent in zip(actions,self.agents):
            ur=torch.as_tensor(agent.u_range,dtype=torch.float32)
            assert not ((a>ur)|(a<-ur)).any(); agent.action.u=a.clone().to(torch.float32)
        for agent in self.world.agents: self.scenario.env_process_action(agent)
        self.scenario.pre_step(); self.world.step(); self.scenario.post_step(); self.steps+=1
        return self.get_from_scenario(True,True,True,True)
def _mod(name,**attrs):
    m=types.ModuleType(name); m.__path__=[]; m.__dict__.update(attrs); sys.modules[name]=m; return m
_mod('vmas',render_interactively=lambda *a,**k:None)
_mod('vmas.simulator'); _mod('vmas.simulator.rendering')
_mod('vmas.simulator.core',Agent=Agent,AgentState=AgentState,EntityState=EntityState,Box=Box,World=World)
_mod('vmas.simulator.scenario',BaseScenario=BaseScenario)
_mod('vmas.simulator.dynamics'); _mod('vmas.simulator.dynamics.common',Dynamics=Dynamics)
_mod('vmas.simulator.utils',TorchUtils=TorchUtils,override=_override,save_video=lambda *a,**k:None,Color=_mk('Color'),ScenarioUtils=_mk('SU'))
_mod('vmas.simulator.environment',Environment=Environment)
sys.path.insert(0,'/root/reference')
